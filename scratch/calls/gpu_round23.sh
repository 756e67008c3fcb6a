#!/bin/bash
# default bench on the final library with the committed traffic table in place (roofline.traffic filled in)
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r2last_bench.json 2> gpurun_out/r2last_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2last_bench.err; head -c 600 gpurun_out/r2last_bench.json
