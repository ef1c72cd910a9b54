"""List the loops (backward branches) of one kernel in an object file with their instruction mix.
    python tools/sass_loops.py <object-or-so> <kernel-name-substring> [min_instructions]"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 8
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in blk.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(f"== {name[:150]}  ({len(ins)} instructions)")
    for i, (addr, text) in enumerate(ins):
        m = re.search(r"BRA.*?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= addr:
            tgt = int(m.group(1), 16)
            body = [t for a, t in ins if tgt <= a <= addr]
            if len(body) < minlen:
                continue
            mix = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] for t in body)
            print(f"  loop {tgt:#x}..{addr:#x}: {len(body)} instr: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
