#!/bin/bash
# Helper kernels of the training step after a change: the training tests, two bench runs, and the kernels' launch times / instruction counts.
mkdir -p gpurun_out/r03m
python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py -m gpu -x -q 2>&1 | tail -1
for r in 1 2; do python bench.py --workload train_step --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"; done
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"head_bwd|maxpool_bwd" -c 5 --csv --log-file gpurun_out/r03m/hb.csv python bench.py --workload train_step --steps 1 --warmup 1 > /dev/null 2>&1
grep -E "head_bwd|maxpool" gpurun_out/r03m/hb.csv | awk -F"\",\"" '{print substr($5,1,28), $(NF-2), $(NF)}'
