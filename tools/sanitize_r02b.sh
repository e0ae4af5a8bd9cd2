#!/bin/bash
# compute-sanitizer over the kernels changed in the second half of round 2 (restructured noise kernel, first-layer staging tile,
# residual epilogue / ResUnet, metric pass).   gpurun --timeout 1500 -- 'bash tools/sanitize_r02b.sh'
set -u
OUT=gpurun_out/r02_sanitize_b
mkdir -p "$OUT"
run() {  # tool, label, timeout, pytest args...
  local tool=$1 label=$2 limit=$3; shift 3
  timeout "$limit" compute-sanitizer --tool "$tool" --error-exitcode 99 --report-api-errors no --print-limit 20 \
      python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > "$OUT/${tool}_${label}.log" 2>&1
  local rc=$?
  echo "$tool $label rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/${tool}_${label}.log" | tail -1) : $(grep -E 'passed|failed|error' "$OUT/${tool}_${label}.log" | tail -1)" | tee -a "$OUT/summary.txt"
}
run memcheck noise 500 tests/test_gpu_noise.py -k "philox_kernel_uses or shard_independence or device_philox_words or tiny_tukey or row_noise_is_constant or tail_refinement"
run racecheck noise 500 tests/test_gpu_noise.py -k "philox_kernel_uses or shard_independence or device_philox_words or tiny_tukey"
run memcheck first_resunet_eval 500 tests/test_gpu_unet.py tests/test_gpu_eval.py -k "(fused_first or resunet or ResUnet or psnr_ssim) and not 1424 and not 512 and not full_frame and not 1744 and not 1736"
run racecheck first_resunet_eval 500 tests/test_gpu_unet.py tests/test_gpu_eval.py -k "(fused_first or resunet or ResUnet or psnr_ssim) and not 1424 and not 512 and not full_frame and not 1744 and not 1736"
cat "$OUT/summary.txt"
