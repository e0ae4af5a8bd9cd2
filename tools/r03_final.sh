#!/bin/bash
# Last verification pass of round 2 (run under gpurun; outputs under gpurun_out/r03j/): the -m gpu suite, smoke, the default bench
# line, the reference arm, the training-step bench and the helper kernels' launch times.
O=gpurun_out/${R03_OUT:-r03j}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > $O/bench_path64.json 2> $O/bench_path64.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload train_step --steps 50 --warmup 5 > $O/bench_train_step.json 2> $O/bench_train_step.err
for f in path64 reference train_step; do python - $O/bench_$f.json $f <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d["ms_per_step"], 4), "ms", round(d["value"], 1), d["unit"], "e2e", round((d.get("e2e") or {}).get("value", 0), 1), (d.get("clocks") or {}).get("sm_mhz"), (d.get("roofline") or {}).get("frac"), ((d.get("roofline_parts") or {}).get("train_step") or {}).get("ms_per_step"))
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_bwd|l1_loss|maxpool_bwd" -c 6 --csv --log-file $O/launches_small.csv python bench.py --workload train_step --steps 1 --warmup 1 > /dev/null 2>&1
grep -E "head_bwd|l1_loss|maxpool" $O/launches_small.csv | awk -F'","' '{print substr($5,1,40), $(NF)}' | head -6
