#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (step kernel, autocast loss, eval rows + NMS filter, dealing modes)
mkdir -p gpurun_out
SEL='one_launch or half_precision or single_use or sharded or kps_to_dict or threshold or ties or dealing or basic_encoder or pipeline_rejects or incremental or loss_kernel_variants or fused_kernel_variants'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_solver_loop.py tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/san2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/san2_$tool.log | tail -3
done
