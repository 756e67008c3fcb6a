#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or train or acc" > gpurun_out/t7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t7_pytest.log
for e in "" "noacc"; do
SP_EXP_TRAIN=$e timeout 300 python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 5 --env SP_TRAIN_TILE_CFG=0 --env SP_TRAIN_TILE_CFG=1 --env SP_TRAIN_TILE_CFG=2 --env SP_TRAIN_TILE_CFG=3 --env SP_TRAIN_TILE_CFG=0,SP_TRAIN_WARPS=12 --env SP_TRAIN_NO_TILE=1 2>&1 | sed "s/^/[$e] /" | tee -a gpurun_out/ub_t7.log
SP_EXP_TRAIN=$e timeout 300 python scratch/ubench.py --ops train_fused --batch 512 --hw 96x72 --reps 5 --env SP_TRAIN_TILE_CFG=0 --env SP_TRAIN_TILE_CFG=1 --env SP_TRAIN_NO_TILE=1 2>&1 | sed "s/^/[$e] /" | tee -a gpurun_out/ub_t7.log
done
