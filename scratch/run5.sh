#!/bin/bash
mkdir -p gpurun_out
for e in "noacc,nograd" "nograd" "noacc" ""; do
SP_EXP_TRAIN=$e timeout 300 python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 5 --env SP_TRAIN_PPC=1 --env SP_TRAIN_PPC=2 --env SP_TRAIN_PPC=4 --env SP_TRAIN_PPC=8 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=3 --env SP_TRAIN_PPC=2,SP_TRAIN_RING=1 --env SP_TRAIN_PPC=1,SP_TRAIN_WARPS=12 --env SP_TRAIN_PPC=1,SP_TRAIN_WARPS=8 2>&1 | sed "s/^/[$e] /" | tee -a gpurun_out/ub_t5.log
done
