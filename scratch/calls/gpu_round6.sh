#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --env "" --env "SP_TRAIN_TILE_CFG=2" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=2" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=4" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=5" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=4,SP_TRAIN_WARPS=12" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=4,SP_TRAIN_WARPS=8"
python scratch/ubench.py --ops train_fused --batch 512 --hw 96x72 --env "" --env "SP_TRAIN_BULK_STORE=1" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=2" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_TILE_CFG=2,SP_TRAIN_WARPS=6" --env "SP_TRAIN_BULK_STORE=1,SP_TRAIN_WARPS=12"
python scratch/ubench.py --ops step --batch 1024,2048 --hw 64x48 --env "" --env "SP_STEP_WARPS=9" --env "SP_STEP_WARPS=10"
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
timeout 200 python scratch/gpu_fuzz.py 60 7 > gpurun_out/${TAG}_fuzz.log 2>&1; tail -5 gpurun_out/${TAG}_fuzz.log
