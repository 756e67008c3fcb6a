#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c0_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c0_pytest.log
( time timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/c0_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c0_smoke.log
( time timeout 400 python bench.py ) > gpurun_out/c0_bench.json 2> gpurun_out/c0_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/c0_bench.err
