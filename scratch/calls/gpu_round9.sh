#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops step --batch 1024,2048,512 --hw 64x48 --env "" --env "SP_STEP_STATIC_PCT=0" --env "SP_STEP_STATIC_PCT=30" --env "SP_STEP_STATIC_PCT=80" --env "SP_STEP_STATIC_PCT=100" --env "SP_STEP_STATIC_PCT=60,SP_STEP_WARPS=7" --env "SP_STEP_STATIC_PCT=60,SP_STEP_WARPS=9"
python scratch/ubench.py --ops step --batch 512 --hw 96x72 --env "" --env "SP_STEP_STATIC_PCT=0" --env "SP_STEP_STATIC_PCT=100" --env "SP_STEP_STATIC_PCT=30" --env "SP_STEP_STATIC_PCT=60,SP_STEP_WARPS=7"
python scratch/ubench.py --ops step --batch 128,256 --hw 64x48 --env "" --env "SP_STEP_STATIC_PCT=50"
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
timeout 200 python scratch/gpu_fuzz.py 60 17 > gpurun_out/${TAG}_fuzz.log 2>&1; tail -4 gpurun_out/${TAG}_fuzz.log
