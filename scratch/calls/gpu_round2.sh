#!/bin/bash
# one gpurun call (1 GPU): GPU tests, bench, eval stage timing
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 1200 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python scratch/eval_stages.py > gpurun_out/${TAG}_eval_stages.log 2>&1; echo "stages rc=$?"; cat gpurun_out/${TAG}_eval_stages.log
timeout 600 python bench.py --steps 50 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
