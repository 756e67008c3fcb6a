#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 600 python -m pytest tests/test_gpu_solver_loop.py -m gpu --maxfail=10 -q -k "one_launch" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
{
python scratch/ubench.py --ops step --batch 1024 --hw 64x48 --env "" --env "SP_STEP_STAGES=1" --env "SP_STEP_WARPS=7" --env "SP_STEP_WARPS=6" --env "SP_STEP_WARPS=5" --env "SP_STEP_WARPS=4" --env "SP_STEP_WARPS=9,SP_STEP_STAGES=1" --env "SP_STEP_WARPS=7,SP_STEP_STAGES=1" --env "SP_STEP_WARPS=6,SP_STEP_STAGES=1"
python scratch/ubench.py --ops step --batch 256,512,4096 --hw 64x48 --env "" --env "SP_STEP_WARPS=8,SP_STEP_STAGES=2" --env "SP_STEP_WARPS=15,SP_STEP_STAGES=1" --env "SP_STEP_WARPS=11,SP_STEP_STAGES=1"
python scratch/ubench.py --ops step --batch 512 --hw 96x72 --env "" --env "SP_STEP_WARPS=4" --env "SP_STEP_WARPS=3,SP_STEP_STAGES=2" --env "SP_STEP_WARPS=4,SP_STEP_STAGES=1"
python scratch/ubench.py --ops loss --batch 512 --hw 96x72 --env "" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=5"
python scratch/ubench.py --ops loss --batch 2048,256 --hw 64x48 --env "" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=4" --env "SP_LOSS_BULK_STORE=1,SP_LOSS_RING=5"
} > gpurun_out/${TAG}_ubench.log 2>&1; echo "ubench rc=$?"; cat gpurun_out/${TAG}_ubench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_|eval_rows|box_affine" -c 60 --csv --log-file gpurun_out/${TAG}_eval_launches.csv python scratch/eval_stages.py > gpurun_out/${TAG}_ncu_eval.log 2>&1; echo "ncu rc=$?"
python - $TAG <<'PY'
import csv,collections,sys
rows=[r for r in csv.reader(open('gpurun_out/'+sys.argv[1]+'_eval_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:]:
    print(r[ki][:60], r[gi], r[vi])
PY
