#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 1200 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python scratch/eval_stages.py > gpurun_out/${TAG}_eval_stages.log 2>&1; echo "stages rc=$?"; cat gpurun_out/${TAG}_eval_stages.log
timeout 900 python bench.py --steps 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
