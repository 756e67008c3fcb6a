#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2p}
{
python scratch/ubench.py --ops step --batch 1024,2048,512 --hw 64x48 --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=9" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=10" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=11" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=12" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=14" --env "SP_STEP_STATIC_PCT=100,SP_STEP_WARPS=10"
python scratch/ubench.py --ops step --batch 512,1024 --hw 96x72 --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=6" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=7" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=5"
python scratch/ubench.py --ops step --batch 256,384 --hw 64x48 --env "" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=10" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=13" --env "SP_STEP_STATIC_PCT=80,SP_STEP_WARPS=15"
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
