#!/bin/bash
# Run on the GPU box (under gpurun): compute-sanitizer over the op-level GPU tests and the ragged-geometry probe.
# memcheck slows kernels 10-50x: the selection below keeps it to a few minutes.  Outputs go to gpurun_out/.
set -u
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20"
$S python -m pytest tests/test_gpu_ops.py -x -q -k "selective_scan or xdt or dwconv or ss2d or sampler or final_conv or ln_ or gn_ or transposed" \
    > gpurun_out/sanitize_ops.log 2>&1; echo "ops: rc=$?" | tee -a gpurun_out/sanitize_ops.log
FD_ALLOW_ODD_16BIT=1 $S python tools/probes/ragged_16bit.py > gpurun_out/sanitize_ragged.log 2>&1; echo "ragged: rc=$?" | tee -a gpurun_out/sanitize_ragged.log
$S python -m pytest tests/test_gpu_model.py -x -q -k "unet_forward or ragged" > gpurun_out/sanitize_model.log 2>&1; echo "model: rc=$?" | tee -a gpurun_out/sanitize_model.log
tail -5 gpurun_out/sanitize_*.log
