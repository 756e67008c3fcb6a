#!/bin/bash
# ncu capture: register stores vs TMA bulk stores in the loss and fused kernels (profiles/prof_store_path.py)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"mse_ring|encode_mse_tile" -f -o /tmp/r2store python profiles/prof_store_path.py > gpurun_out/r2store_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2store_ncu.log
ncu -i /tmp/r2store.ncu-rep --page raw --csv > gpurun_out/r2store_raw.csv 2>/dev/null; wc -l gpurun_out/r2store_raw.csv
