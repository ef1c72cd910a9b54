"""Static SASS evidence per kernel of libfounddiff_b200.so: counts of the mnemonics that prove tcgen05 / TMEM / TMA / mma.sync /
cp.async use (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP) plus the
instruction total, so that a reader can check which kernels are tensor-core / TMA kernels without a GPU.
    python tools/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "founddiff_b200", "libfounddiff_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = [("UTC*MMA", r"UTC\w*MMA"), ("LDTM", r"LDTM"), ("STTM", r"STTM"), ("UTMALDG", r"UTMALDG"), ("UTMASTG", r"UTMASTG"),
        ("SYNCS", r"SYNCS"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"LDGSTS"), ("LDSM", r"LDSM"), ("MUFU", r"MUFU"), ("SHFL", r"SHFL"),
        ("FFMA", r"\bFFMA(?!2)"), ("F*2 f32x2", r"\bF(?:FMA|MUL|ADD)2")]
cur, rows = None, collections.OrderedDict()
ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    m = ins.match(line)
    if m and cur:
        op = m.group(1)
        rows[cur]["total"] += 1
        for name, pat in KEYS:
            if re.match(pat, op):
                rows[cur][name] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


agg = collections.OrderedDict()
for n, c in rows.items():
    d = demangle(n)
    base = re.sub(r"^void ", "", d)
    base = re.sub(r"\(anonymous namespace\)::", "", base)
    base = re.split(r"[<(]", base)[0]
    a = agg.setdefault(base, collections.Counter())
    a["variants"] += 1
    for k, v in c.items():
        a[k] = max(a[k], v)          # per kernel template: the maximum over its instantiations
print(f"# {os.path.relpath(lib, ROOT)}: static SASS mnemonic counts per kernel template (maximum over its instantiations), sm_100a")
hdr = ["variants", "total"] + [k for k, _ in KEYS]
print(f"{'kernel':34s} " + " ".join(f"{h:>8s}" for h in hdr))
for base, c in sorted(agg.items(), key=lambda kv: (-kv[1]["UTC*MMA"], -kv[1]["HMMA"], kv[0])):
    print(f"{base[:34]:34s} " + " ".join(f"{c[h]:8d}" for h in hdr))
