#!/bin/bash
# Third session: compute-sanitizer over the kernels added in it (chained time-sliced scan, tcgen05 flash attention).  Run under gpurun.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  C="compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20"
  timeout -s KILL 400 $C python -m pytest tests/test_gpu_ss2d_tm.py -x -q -k "chained" > gpurun_out/s3_${tool}_chain.log 2>&1; echo "chained scan $tool: rc=$?" | tee -a gpurun_out/s3_${tool}_chain.log
  timeout -s KILL 400 $C python -m pytest tests/test_gpu_ops.py -x -q -k "flash_attention and tc" > gpurun_out/s3_${tool}_flash.log 2>&1; echo "flash_attn_d32_tc $tool: rc=$?" | tee -a gpurun_out/s3_${tool}_flash.log
done
grep -h "rc=\|ERROR SUMMARY\|passed\|failed" gpurun_out/s3_*.log
