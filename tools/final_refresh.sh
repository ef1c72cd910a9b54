#!/bin/bash
# End-of-round evidence on the GPU box (under gpurun): smoke, the driver-style bench line + kernel table, the micro-benchmarks,
# then the ncu launch list and full captures (tools/refresh_profiles.sh).  Post-process here into profiles/ (see profiles/README.md).
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --kernel-table gpurun_out/r2_kernels_final.json > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
timeout 600 python bench_micro.py > gpurun_out/r2_micro_final.txt 2>&1
bash tools/refresh_profiles.sh > gpurun_out/refresh5.log 2>&1
tail -3 gpurun_out/r2_smoke.log; cut -c1-300 gpurun_out/r2_bench_n1_final.json; tail -5 gpurun_out/refresh5.log
