#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of the bench command + `--set full` captures of the dominant kernels.
# Outputs go to gpurun_out/; summarise them here with tools/ncu_summary.py / tools/launch_list_summary.py into profiles/.
set -u
mkdir -p gpurun_out
# 1. launch list: per-launch durations over ~2 sample() calls of the default workload (cold cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1500 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-profile --no-cpu-baseline --no-extras > gpurun_out/final_launches.log 2>&1
# 2. full captures
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tw2 -c 1 -o gpurun_out/final_scan_tw2 -f \
    python tools/prof_scan_tm.py chain > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -o gpurun_out/final_conv_3x3 -f \
    python bench_micro.py --only conv --pick 3 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -o gpurun_out/final_conv_inproj -f \
    python bench_micro.py --only conv --pick 0 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_gate_out_proj -c 1 -o gpurun_out/final_tail -f \
    python tools/bench_tail.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_tm -c 1 -o gpurun_out/final_dwconv_tm -f \
    python bench_micro.py --only tm --iters 1 > /dev/null 2>&1
ls -la gpurun_out/final_* | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_d32_tc -c 1 -o gpurun_out/final_flash_tc -f \
    python bench_micro.py --only flash --pick 1 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/final_* | head -20
