#!/bin/bash
# 1 GPU: smoke, GPU tests, bench (both arms), ncu launch list + full capture reduced to CSV
mkdir -p gpurun_out
TAG=${1:-r2final}
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
SP_BENCH_CYCLES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_|mse_|decode_|step_|oks_|rescore_|pack_|heatmap_acc|scale_inplace|train_geometry|box_affine|eval_rows" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-ops --no-e2e --no-cpu --no-eval > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launchlist rc=$?"
PROF_REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"decode_|encode_|mse_|oks_|rescore|train_|acc_|step_|eval_rows|box_affine" -f -o /tmp/${TAG}_prof python profiles/prof_driver.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/${TAG}_ncu_full.log
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv -k regex:"step_kernel" > gpurun_out/${TAG}_source_step.csv 2>/dev/null
SZ=$(stat -c %s /tmp/${TAG}_prof.ncu-rep); echo "report bytes $SZ"
ls -la gpurun_out | grep ${TAG}; du -sh gpurun_out
