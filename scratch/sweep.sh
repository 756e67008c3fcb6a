python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ENVS="--env SP_TRAIN_WARPS=16"
for w in 12 14 16; do for c in 96 192 384; do for r in 1 2 3; do ENVS="$ENVS --env SP_TRAIN_WARPS=$w,SP_TRAIN_CHUNK_QUADS=$c,SP_TRAIN_RING=$r"; done; done; done
python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 5 $ENVS 2>&1 | tee gpurun_out/ub6.log | sort -k9 -n | head -12
python scratch/ubench.py --ops train_fused,loss --batch 512,2048 --hw 96x72 --reps 5 2>&1 | tee -a gpurun_out/ub6.log
