#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c1_pytest.log
( time timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke ok" gpurun_out/c1_smoke.log
( time timeout 400 python bench.py ) > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/c1_bench.err
