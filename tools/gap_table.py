"""Per kernel class: measured time per sample() call vs the time the same launches would take at the measured roofs
(max(algorithmic bytes / HBM peak, algorithmic FLOPs / tensor peak) per launch, from the `bench.py --kernel-table` JSON).
    python tools/gap_table.py profiles/r1_kernels_final.json > profiles/r1_gap_by_kernel_class.txt"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6546.2))
tc = float(peaks.get("bf16_tflops", 1656.0))
d = json.load(open(sys.argv[1]))
agg = collections.OrderedDict()
for k in d["kernels"]:
    t_h = k["alg_GB"] / hbm * 1e3
    t_t = k["alg_GFLOP"] / tc if k["kernel"].startswith("conv") or "tc" in k["kernel"] else 0.0
    roof = max(t_h, t_t) * k["launches"]
    name = k["kernel"]
    if name == "selective_scan_merge":
        name += " (channel-per-lane)" if k["shape"].endswith("cl") else " (warp-shuffle)"
    a = agg.setdefault(name, [0.0, 0.0, 0])
    a[0] += k["total_ms"]
    a[1] += roof if (k["alg_GB"] or k["alg_GFLOP"]) else float("nan")
    a[2] += k["launches"]
tot = sum(v[0] for v in agg.values())
print(f"# {os.path.basename(sys.argv[1])}: {tot:.2f} ms per sample() call (eager replay, CUDA events per launch); roofs: HBM {hbm:.0f} GB/s, bf16 {tc:.0f} TFLOP/s")
print("# roof = sum over launches of max(algorithmic bytes / HBM, algorithmic FLOPs / tensor); addends / gates of fused epilogues are NOT counted")
print(f"{'kernel class':44s} {'launches':>8s} {'ms':>8s} {'share':>6s} {'roof ms':>8s} {'roof/ms':>8s} {'excess ms':>9s}")
for name, (ms, roof, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    ok = roof == roof and roof > 0
    print(f"{name:44s} {n:8d} {ms:8.2f} {ms / tot:6.3f} {roof if ok else float('nan'):8.2f} {roof / ms if ok else float('nan'):8.2f} {ms - roof if ok else float('nan'):9.2f}")
