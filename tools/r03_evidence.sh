#!/bin/bash
# Evidence pass at the end of round 2 (run under gpurun; outputs under gpurun_out/r03g/): the launch list of the training-step
# bench, and ncu --set full captures of the weight-gradient kernel and of the fused first layer in their final r02 form.
O=gpurun_out/r03g; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_train_step.csv \
  python bench.py --workload train_step --steps 2 --warmup 1 > $O/launches_train.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_nhwc -c 25 -o $O/prof_wgrad -f \
  python bench.py --workload train_step --steps 1 --warmup 1 > $O/ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_first -c 1 -o $O/prof_first -f \
  python tools/prof_unet64.py > $O/ncu_first.log 2>&1
ls -la $O
