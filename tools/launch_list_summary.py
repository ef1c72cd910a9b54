"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
usage: python tools/launch_list_summary.py gpurun_out/final_launches.csv > profiles/<name>.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    rows.append((r[ik], us))
OURS = ("conv_tc", "selective_scan", "scan_tw", "scan_tm", "row_rstd", "dwconv", "ln_", "gn_", "x_proj", "xdt_proj", "init_conv", "gram_mma", "attn_weff", "final_conv",
        "linear_small", "avgpool2x2_nhwc", "sampler_init", "time_sinusoid", "unnormalize", "conv_simt", "merge_", "flash", "linattn")


def base(name):
    n = name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"[<(].*", "", n)
    return n.strip()


agg = collections.OrderedDict()
for k, us in rows:
    b = base(k)
    a = agg.setdefault(b, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if any(o in k for o in OURS))
print(f"# {len(rows)} launches, total {total:.0f} us (cold-cache, serialised launches: compare SHARES, not absolutes)")
print(f"# hand-written founddiff_b200 kernels: {ours / total:.3f} of the device time")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    tag = "[ours]" if any(o in k for o in OURS) else "[torch/cuDNN: DA-CLIP stem + attention pool, memsets, copies]"
    print(f"{k[:62]:62s} n={n:4d} total_us={us:10.1f} share={us / total:.3f} {tag}")
