#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "encode" > gpurun_out/t26_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t26_pytest.log
timeout 400 python scratch/ubench.py --ops encode --batch 512,1024,128 --hw 96x72 --reps 9 --env SP_ENCODE_PARTS=1 --env "" --env SP_ENCODE_PARTS=2 --env SP_ENCODE_PARTS=3 --env SP_ENCODE_PARTS=4 2>&1 | tee gpurun_out/ub26.log
timeout 400 python scratch/ubench.py --ops encode --batch 1024,2048,128 --hw 64x48 --reps 9 --env SP_ENCODE_PARTS=1 --env "" --env SP_ENCODE_PARTS=2 --env SP_ENCODE_PARTS=4 2>&1 | tee -a gpurun_out/ub26.log
