#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or train or acc" > gpurun_out/t2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t2_pytest.log
ENVS="--env SP_TRAIN_NO_TILE=1 --env SP_TRAIN_PPC=1 --env SP_TRAIN_PPC=2 --env SP_TRAIN_PPC=4 --env SP_TRAIN_PPC=2,SP_TRAIN_RING=3 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=4 --env SP_TRAIN_PPC=2,SP_TRAIN_WARPS=12,SP_TRAIN_RING=3 --env SP_TRAIN_PPC=4,SP_TRAIN_WARPS=8 --env SP_TRAIN_PPC=2,SP_TRAIN_WARPS=8,SP_TRAIN_RING=4"
timeout 600 python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 7 $ENVS 2>&1 | tee gpurun_out/ub_t2.log
ENVS="--env SP_TRAIN_NO_TILE=1 --env SP_TRAIN_PPC=1 --env SP_TRAIN_PPC=2 --env SP_TRAIN_PPC=1,SP_TRAIN_RING=3 --env SP_TRAIN_PPC=1,SP_TRAIN_WARPS=12,SP_TRAIN_RING=3 --env SP_TRAIN_PPC=1,SP_TRAIN_WARPS=8,SP_TRAIN_RING=3"
timeout 600 python scratch/ubench.py --ops train_fused --batch 512 --hw 96x72 --reps 7 $ENVS 2>&1 | tee -a gpurun_out/ub_t2.log
timeout 300 python scratch/ubench.py --ops encode,loss,train_fused,decode,flip_decode --batch 128,256,512 --hw 64x48 --reps 7 2>&1 | tee -a gpurun_out/ub_t2.log
