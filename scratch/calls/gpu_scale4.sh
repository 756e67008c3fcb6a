#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 --no-ops --no-e2e --no-cpu > gpurun_out/r2x_bench_4gpu.json 2> gpurun_out/r2x_bench_4gpu.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2x_bench_4gpu.json
