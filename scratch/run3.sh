#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or train or acc" > gpurun_out/t3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t3_pytest.log
for e in "" noacc nograd "noacc,nograd"; do
SP_EXP_TRAIN=$e timeout 300 python scratch/ubench.py --ops train_fused --batch 1024 --hw 64x48 --reps 7 --env SP_TRAIN_PPC=1 --env SP_TRAIN_NO_TILE=1 2>&1 | sed "s/^/[$e] /" | tee -a gpurun_out/ub_t3.log
done
timeout 600 ./scratch/stream_bench 256 2>&1 | tee gpurun_out/stream_bench.log | grep -E "memcpy|reg|nread=1" 
