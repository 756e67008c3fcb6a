#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 5 --env SP_TRAIN_PPC=2,SP_TRAIN_RING=1 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=1,SP_TRAIN_WARPS=32 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=2,SP_TRAIN_WARPS=32 --env SP_TRAIN_PPC=2,SP_TRAIN_RING=1,SP_TRAIN_WARPS=32  --env SP_TRAIN_PPC=2,SP_TRAIN_RING=1,SP_TRAIN_WARPS=24 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=2,SP_TRAIN_WARPS=24 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=2,SP_TRAIN_WARPS=20 2>&1 | tee -a gpurun_out/ub_t6.log
