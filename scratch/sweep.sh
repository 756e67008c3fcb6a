python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ENVS="--env SP_LOSS_FORCE_LDG=1"
for w in 4 6 8 12 16; do for c in 96 192 384; do for r in 2 3 4; do ENVS="$ENVS --env SP_LOSS_WARPS=$w,SP_LOSS_CHUNK_QUADS=$c,SP_LOSS_RING=$r"; done; done; done
python scratch/ubench.py --ops loss --batch 1024,4096 --hw 64x48 --reps 5 $ENVS 2>&1 | tee gpurun_out/ub4.log | sort -k9 -n | head -70
