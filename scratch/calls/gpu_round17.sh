#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scratch/gpu_fuzz.py 150 23 > gpurun_out/r2fuzz.log 2>&1; tail -6 gpurun_out/r2fuzz.log
