#!/bin/bash
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool initcheck --error-exitcode 86 --print-limit 3 python -m pytest tests/test_gpu_solver_loop.py tests/test_gpu_parity.py -m gpu -x -v -k "one_launch or half_precision or single_use or sharded or kps_to_dict or threshold or ties or dealing or basic_encoder or pipeline_rejects or incremental or loss_kernel_variants or fused_kernel_variants" 2>&1 | grep -E "PASSED|FAILED|Uninitialized|     at |ERROR SUMMARY|passed|failed" | cut -c1-220 | head -150 > gpurun_out/initcheck_probe.log
tail -40 gpurun_out/initcheck_probe.log
