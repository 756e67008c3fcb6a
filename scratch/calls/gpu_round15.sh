#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2y}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver_loop.py -m gpu --maxfail=10 -q -k "loss or step or solver" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops loss --batch 1024,2048,512 --hw 64x48 --env "SP_LOSS_TAIL_PCT=0" --env "SP_LOSS_TAIL_PCT=4" --env "SP_LOSS_TAIL_PCT=8" --env "SP_LOSS_TAIL_PCT=12" --env "SP_LOSS_TAIL_PCT=20" --env "SP_LOSS_TAIL_PCT=8,SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4"
python scratch/ubench.py --ops loss --batch 512,1024 --hw 96x72 --env "SP_LOSS_TAIL_PCT=0" --env "SP_LOSS_TAIL_PCT=8" --env "SP_LOSS_TAIL_PCT=12" --env "SP_LOSS_TAIL_PCT=12,SP_LOSS_BULK_STORE=1,SP_LOSS_RING=5"
python scratch/ubench.py --ops loss --batch 256 --hw 64x48 --env "" --env "SP_LOSS_TAIL_PCT=8,SP_LOSS_TAIL_MIN=1"
} > gpurun_out/${TAG}_ubench.log 2>&1; cat gpurun_out/${TAG}_ubench.log
