#!/bin/bash
# A/B of the training step on ONE box: bias sums of conv{1..4}_2 inside the pool backward (default) against the separate passes.
O=gpurun_out/${AB_OUT:-r03i}; mkdir -p $O
python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_ddp.py -m gpu -x -q > $O/pytest_train.log 2>&1; tail -2 $O/pytest_train.log
for r in 1 2; do
  for v in 1 0; do
    PNNP_POOL_BIAS=$v python bench.py --workload train_step --steps 50 --warmup 5 > $O/train_pb${v}_$r.json 2> $O/train_pb${v}_$r.err
    python - $O/train_pb${v}_$r.json $v <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("pool_bias", sys.argv[2], round(d["ms_per_step"], 4), "ms", d["clocks"]["sm_mhz"])
PY
  done
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_bwd|l1_loss|maxpool_bwd" -c 12 --csv --log-file $O/launches_small.csv python bench.py --workload train_step --steps 1 --warmup 1 > /dev/null 2>&1
grep -E "head_bwd|l1_loss|maxpool" $O/launches_small.csv | awk -F'","' '{print substr($5,1,40), $(NF)}' | head -12
