#!/bin/bash
# one gpurun call (1 GPU): GPU tests, micro-benchmarks of the new kernels, eval stage timing, bench
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 1200 python -m pytest tests -m gpu --maxfail=10 -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops loss --batch 1024 --hw 64x48,96x72 --env "" --env "SP_LOSS_BULK_STORE=1" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4,SP_LOSS_WARPS=3" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=2,SP_LOSS_WARPS=6" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=5" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4,SP_LOSS_CHUNK_QUADS=384"
python scratch/ubench.py --ops step --batch 128,1024 --hw 64x48 --env "" --env "SP_STEP_WARPS=12" --env "SP_STEP_WARPS=10" --env "SP_STEP_WARPS=15" --env "SP_STEP_WARPS=8"
python scratch/ubench.py --ops step --batch 512 --hw 96x72 --env "" --env "SP_STEP_WARPS=6" --env "SP_STEP_WARPS=5"
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
timeout 300 python scratch/eval_stages.py > gpurun_out/${TAG}_eval_stages.log 2>&1; echo "stages rc=$?"; cat gpurun_out/${TAG}_eval_stages.log
timeout 600 python bench.py --steps 50 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
