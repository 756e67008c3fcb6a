#!/bin/bash
mkdir -p gpurun_out
SP_TRAIN_PPC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode_mse_tile" -c 2 -f -o gpurun_out/tile_prof python scratch/prof_one.py train_fused 64 48 1024 > gpurun_out/tile_prof.log 2>&1; echo rc=$?
SP_EXP_TRAIN=noacc,nograd SP_TRAIN_PPC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode_mse_tile" -c 2 -f -o gpurun_out/tile_prof_ro python scratch/prof_one.py train_fused 64 48 1024 > gpurun_out/tile_prof_ro.log 2>&1; echo rc=$?
