#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 7 --env SP_TRAIN_WARPS=16 --env SP_TRAIN_WARPS=15 --env SP_TRAIN_WARPS=14 --env SP_TRAIN_WARPS=13 --env SP_TRAIN_WARPS=12 --env SP_TRAIN_WARPS=10 --env SP_TRAIN_WARPS=8 2>&1 | tee gpurun_out/ub8.log
timeout 300 python scratch/ubench.py --ops train_fused --batch 512 --hw 96x72 --reps 7 --env SP_TRAIN_WARPS=16 --env SP_TRAIN_WARPS=15 --env SP_TRAIN_WARPS=14 --env SP_TRAIN_WARPS=13 --env SP_TRAIN_WARPS=12 --env SP_TRAIN_WARPS=10 --env SP_TRAIN_WARPS=8 2>&1 | tee -a gpurun_out/ub8.log
timeout 300 python scratch/ubench.py --ops train_fused --batch 1000,2048 --hw 64x48 --reps 7 --env SP_TRAIN_WARPS=16 --env SP_TRAIN_WARPS=15 --env SP_TRAIN_WARPS=12 2>&1 | tee -a gpurun_out/ub8.log
timeout 300 python scratch/ubench.py --ops decode --batch 1024 --hw 64x48 --reps 7 --env SP_DECODE_WARPS=16 --env SP_DECODE_WARPS=15 --env SP_DECODE_WARPS=12 --env SP_DECODE_WARPS=8,SP_DECODE_STAGES=2 2>&1 | tee -a gpurun_out/ub8.log
