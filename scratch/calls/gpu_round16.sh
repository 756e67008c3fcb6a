#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/eval_stages.py > gpurun_out/r2z_eval_stages.log 2>&1; cat gpurun_out/r2z_eval_stages.log
bash scratch/gpu_final.sh r2final
