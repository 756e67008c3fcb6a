#!/bin/bash
mkdir -p gpurun_out
PROF_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode_mse_tile|decode_tma" -f -o gpurun_out/x10_prof python profiles/prof_driver.py > gpurun_out/x10_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/x10_ncu.log; ls -la gpurun_out/
