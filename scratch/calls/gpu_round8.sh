#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 200 python scratch/gpu_fuzz.py 70 13 > gpurun_out/${TAG}_fuzz.log 2>&1; tail -4 gpurun_out/${TAG}_fuzz.log
