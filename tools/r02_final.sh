#!/bin/bash
# Final measurement pass of round 2 (run under gpurun; outputs under gpurun_out/r02_final/)
O=gpurun_out/${R02_OUT:-r02_final}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_path64.json 2> $O/bench_path64.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
for w in synth64 unet_sony train_step imx686_eval sony_evaltest; do
  python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
python bench.py --workload unet_sony --precision tf32 --steps 50 --warmup 5 > $O/bench_unet_sony_tf32.json 2> $O/bench_unet_sony_tf32.err
python tools/profile_unet.py 1,4,1424,2128 > $O/unet_layer_times_sony.txt 2>&1
python tools/profile_unet.py 64,4,512,512 > $O/unet_layer_times_64crops.txt 2>&1
python tools/profile_unet.py 1,4,1424,2128 --resunet > $O/resunet_layer_times_sony.txt 2>&1
python tools/synth_exp.py > $O/synth_exp.txt 2>&1
python tools/e2e_split.py > $O/e2e_split.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_path64.csv python bench.py --steps 2 --warmup 1 > $O/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:noise_synth_fast -s 2 -c 1 -o $O/prof_synth -f python tools/prof_synth.py > $O/ncu_synth.log 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], round(d.get("ms_per_step", 0), 3), "ms", round(d.get("value", 0), 1), d.get("unit"), "e2e", round((d.get("e2e") or {}).get("value", 0), 1),
          "frac", (d.get("roofline") or {}).get("frac"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("cpu_baseline") or {}).get("kind"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
