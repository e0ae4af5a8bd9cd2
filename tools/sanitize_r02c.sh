#!/bin/bash
# compute-sanitizer over the kernels changed at the end of round 2 (pack with the reciprocal form, metric pass with four quantities /
# 128-thread blocks), plus one ncu --set full capture of the metric pass.   gpurun --timeout 900 -- 'bash tools/sanitize_r02c.sh'
set -u
OUT=gpurun_out/r02_sanitize_c
mkdir -p "$OUT"
run() {  # tool, label, timeout, pytest args...
  local tool=$1 label=$2 limit=$3; shift 3
  timeout "$limit" compute-sanitizer --tool "$tool" --error-exitcode 99 --report-api-errors no --print-limit 20 \
      python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > "$OUT/${tool}_${label}.log" 2>&1
  local rc=$?
  echo "$tool $label rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/${tool}_${label}.log" | tail -1) : $(grep -E 'passed|failed|error' "$OUT/${tool}_${label}.log" | tail -1)" | tee -a "$OUT/summary.txt"
}
run memcheck pack_eval 400 tests/test_gpu_pack.py tests/test_gpu_eval.py -k "not full_frame and not 1424 and not 2848"
run racecheck pack_eval 400 tests/test_gpu_pack.py tests/test_gpu_eval.py -k "not full_frame and not 1424 and not 2848"
cat "$OUT/summary.txt"
ncu --set full --clock-control none --import-source on -k regex:ssim_mse_v2 --profile-from-start off -c 1 -o $OUT/prof_eval -f python tools/prof_eval.py > $OUT/ncu_eval.log 2>&1
tail -2 $OUT/ncu_eval.log
