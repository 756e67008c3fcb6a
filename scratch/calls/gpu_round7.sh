#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops step --batch 1024,2048,512 --hw 64x48 --env "" --env "SP_STEP_NO_TILE=1" --env "SP_STEP_WARPS=9" --env "SP_STEP_WARPS=7" --env "SP_STEP_WARPS=6"
python scratch/ubench.py --ops step --batch 512,1024 --hw 96x72 --env "" --env "SP_STEP_NO_TILE=1" --env "SP_STEP_WARPS=5" --env "SP_STEP_WARPS=7"
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
timeout 200 python scratch/gpu_fuzz.py 60 11 > gpurun_out/${TAG}_fuzz.log 2>&1; tail -4 gpurun_out/${TAG}_fuzz.log
