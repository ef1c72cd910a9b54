#!/bin/bash
# Run on the GPU box (under gpurun): compute-sanitizer over the op-level GPU tests, the time-major SS2D kernels, the
# ragged-geometry probe and the metrics kernel at the sizes ADVICE r1 flagged.  memcheck slows kernels 10-50x: the selection
# below keeps it to a few minutes.  racecheck runs on the shared-memory kernels.  Outputs go to gpurun_out/.
set -u
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20"
R="compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20"
$S python -m pytest tests/test_gpu_ops.py -x -q -k "selective_scan or xdt or dwconv or ss2d or sampler or final_conv or ln_ or gn_ or groupnorm or transposed" \
    > gpurun_out/sanitize_ops.log 2>&1; echo "ops memcheck: rc=$?" | tee -a gpurun_out/sanitize_ops.log
$S python -m pytest tests/test_gpu_ss2d_tm.py -x -q -k "not segments_are_exact" > gpurun_out/sanitize_tm.log 2>&1; echo "time-major memcheck: rc=$?" | tee -a gpurun_out/sanitize_tm.log
$R python -m pytest tests/test_gpu_ss2d_tm.py -x -q -k "scan_time_major_vs_oracle or pipeline_vs_oracle" > gpurun_out/racecheck_tm.log 2>&1; echo "time-major racecheck: rc=$?" | tee -a gpurun_out/racecheck_tm.log
FD_ALLOW_ODD_16BIT=1 $S python tools/probes/ragged_16bit.py > gpurun_out/sanitize_ragged.log 2>&1; echo "ragged memcheck: rc=$?" | tee -a gpurun_out/sanitize_ragged.log
$S python -m pytest tests/test_gpu_model.py -x -q -k "unet_forward or ragged or batch_composition" > gpurun_out/sanitize_model.log 2>&1; echo "model memcheck: rc=$?" | tee -a gpurun_out/sanitize_model.log
$S python tools/probes/metrics_small.py > gpurun_out/sanitize_metrics.log 2>&1; echo "metrics memcheck: rc=$?" | tee -a gpurun_out/sanitize_metrics.log
grep -h "rc=\|ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_*.log gpurun_out/racecheck_*.log
