#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or train or acc" > gpurun_out/t9_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t9_pytest.log
timeout 300 python scratch/ubench.py --ops train_fused --batch 1024,2048 --hw 64x48 --reps 7 --env SP_TRAIN_TAIL=0 --env "" --env SP_TRAIN_TAIL=8 --env SP_TRAIN_TAIL=12 --env SP_TRAIN_TAIL=16,SP_TRAIN_TILE_CFG=2 2>&1 | tee gpurun_out/ub9.log
timeout 300 python scratch/ubench.py --ops train_fused --batch 512,1024 --hw 96x72 --reps 7 --env SP_TRAIN_TAIL=0 --env "" --env SP_TRAIN_TAIL=8 --env SP_TRAIN_TAIL=12 --env SP_TRAIN_TAIL=16,SP_TRAIN_TILE_CFG=1 2>&1 | tee -a gpurun_out/ub9.log
