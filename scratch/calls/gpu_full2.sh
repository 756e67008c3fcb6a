#!/bin/bash
mkdir -p gpurun_out
TAG=$1; N=$2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_fullbench_${N}gpu.json 2> gpurun_out/${TAG}_fullbench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_fullbench_${N}gpu.err; head -c 1500 gpurun_out/${TAG}_fullbench_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/${TAG}_fullref_${N}gpu.json 2> gpurun_out/${TAG}_fullref_${N}gpu.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_fullref_${N}gpu.json | head -c 400
