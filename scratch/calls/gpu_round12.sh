#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2r}
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops loss --batch 1024,2048 --hw 64x48 --env "" --env "SP_LOSS_INTERLEAVE=1" --env "SP_LOSS_INTERLEAVE=2" --env "SP_LOSS_INTERLEAVE=3" --env "SP_LOSS_INTERLEAVE=6" --env "SP_LOSS_INTERLEAVE=12" --env "SP_LOSS_INTERLEAVE=3,SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4"
python scratch/ubench.py --ops loss --batch 512 --hw 96x72 --env "" --env "SP_LOSS_INTERLEAVE=1" --env "SP_LOSS_INTERLEAVE=3" --env "SP_LOSS_INTERLEAVE=9"
python scratch/ubench.py --ops loss --batch 128,256 --hw 64x48 --env "" --env "SP_LOSS_INTERLEAVE=1" --env "SP_LOSS_INTERLEAVE=3"
python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --env "" --env "SP_TRAIN_STATIC_PCT=70" --env "SP_TRAIN_STATIC_PCT=85" --env "SP_TRAIN_STATIC_PCT=100" --env "SP_TRAIN_STATIC_PCT=30"
python scratch/ubench.py --ops train_fused --batch 512 --hw 96x72 --env "" --env "SP_TRAIN_STATIC_PCT=70" --env "SP_TRAIN_STATIC_PCT=85" --env "SP_TRAIN_STATIC_PCT=100"
python scratch/ubench.py --ops decode,flip_decode --batch 1024,256 --hw 64x48 --env ""
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
