#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (small shapes only; the full-size tests are skipped)
mkdir -p gpurun_out
SEL='not cfg4 and not cfg5 and not multigpu and not back_to_back'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/san_$tool.log | tail -3
done
