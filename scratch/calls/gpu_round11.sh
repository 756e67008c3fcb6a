#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver_loop.py -m gpu --maxfail=10 -q -k "decode or step" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops decode,flip_decode --batch 1024,4096 --hw 64x48 --env "" --env "SP_DECODE_GRID_WIDE=1" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=50" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=80" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=90" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=100"
python scratch/ubench.py --ops decode,flip_decode --batch 512 --hw 96x72 --env "" --env "SP_DECODE_GRID_WIDE=1" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=50" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=80" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=100"
python scratch/ubench.py --ops decode --batch 128,256 --hw 64x48 --env "" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=80" --env "SP_DECODE_GRID_WIDE=1,SP_DECODE_STATIC_PCT=100"
python scratch/ubench.py --ops step --batch 1024,256,128 --hw 64x48 --env ""
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
