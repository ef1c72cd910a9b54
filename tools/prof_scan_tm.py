"""One launch of each time-major scan variant at the level-0 shape, for ncu (tools/bench_scan_tm.py is the timing script)."""
import sys
sys.path.insert(0, ".")
import tools.bench_scan_tm as b  # noqa: E402
variants = [int(v) for v in sys.argv[1:]] or [16, -8]
b.run(16, 512, 128, 4, 4, variants, iters=1)
