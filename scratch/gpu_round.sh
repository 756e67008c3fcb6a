#!/bin/bash
# one gpurun call: GPU tests, bench (ours + reference arm), ncu launch list, ncu full capture
mkdir -p gpurun_out
TAG=${1:-r1b}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_|mse_|decode_|oks_|rescore_|pack_|heatmap_acc|scale_inplace|train_geometry|box_affine" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-ops --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_|encode_|mse_|oks_|rescore|train_|acc_" -f -o gpurun_out/${TAG}_prof python profiles/prof_driver.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out
