#!/bin/bash
# final-library check at 2 GPUs: NCCL/fan-out parity test + default bench at N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rs -rP > gpurun_out/r2fin_multigpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2fin_multigpu_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2fin_bench_2gpu.json 2> gpurun_out/r2fin_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2fin_bench_2gpu.err; head -c 1500 gpurun_out/r2fin_bench_2gpu.json
