#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "oks or nms or eval or cfg" > gpurun_out/t11_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t11_pytest.log
echo "--- pair matrix"; timeout 300 python scratch/eval_stages.py 2>&1 | tee gpurun_out/stages11.log
echo "--- serial loop"; SP_NMS_SERIAL=1 timeout 300 python scratch/eval_stages.py 2>&1 | grep oks_nms | tee -a gpurun_out/stages11.log
