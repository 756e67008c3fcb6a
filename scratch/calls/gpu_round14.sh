#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2u}
{
python scratch/ubench.py --ops train_fused,step_train --batch 1024,2048 --hw 64x48 --env ""
python scratch/ubench.py --ops train_fused,step_train --batch 512 --hw 96x72 --env ""
python scratch/ubench.py --ops train_fused,step_train --batch 128,256 --hw 64x48 --env ""
} > gpurun_out/${TAG}_ubench.log 2>&1; cat gpurun_out/${TAG}_ubench.log
timeout 300 python -m pytest tests/test_gpu_solver_loop.py tests/test_gpu_parity.py -m gpu -q -k "round2_golden or incremental or one_launch" 2>&1 | tail -3
