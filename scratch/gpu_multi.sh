#!/bin/bash
# multi-GPU call: NCCL parity test of the sharded evaluator + bench at N GPUs (eval job incl. matches_single_device)
mkdir -p gpurun_out
TAG=$1; N=$2; EXTRA="${3:---no-ops --no-e2e --no-cpu}"
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rs -rP > gpurun_out/${TAG}_multigpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_multigpu_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_${N}gpu.err; cat gpurun_out/${TAG}_bench_${N}gpu.json | head -c 3000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scratch/eval_multi.py 1,2 > gpurun_out/${TAG}_eval_${N}gpu.log 2> gpurun_out/${TAG}_eval_${N}gpu.err; echo "eval_multi rc=$?"; cat gpurun_out/${TAG}_eval_${N}gpu.log
