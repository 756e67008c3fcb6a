#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2s}
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python scratch/eval_stages.py > gpurun_out/${TAG}_eval_stages.log 2>&1; echo "stages rc=$?"; cat gpurun_out/${TAG}_eval_stages.log
{
python scratch/ubench.py --ops decode,flip_decode,step --batch 1024,4096 --hw 64x48 --env ""
python scratch/ubench.py --ops decode,flip_decode,step --batch 512 --hw 96x72 --env ""
} > gpurun_out/${TAG}_ubench.log 2>&1; cat gpurun_out/${TAG}_ubench.log
bash scratch/sanitize2.sh
