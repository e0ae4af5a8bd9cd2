#!/bin/bash
# compute-sanitizer passes over the -m gpu suite ("race detection / sanitizers: none in the reference; new build:
# compute-sanitizer memcheck + racecheck on the kernels").   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# memcheck: every GPU test file at its small sizes (the BASELINE-size cases are deselected: the tool slows kernels 10-100x);
# racecheck: the kernels that communicate through shared memory (noise synthesis queue, eval reductions, training helpers, conv / wgrad
# pipelines).  Summaries land in gpurun_out/r02_sanitize/ and are copied to profiles/ by hand.
set -u
OUT=gpurun_out/r02_sanitize
mkdir -p "$OUT"
SEL='not scale and not full_frame and not 512 and not 1424 and not baseline and not train_mode and not entry_point and not learns and not graph'
run() {  # tool, label, timeout, pytest args...
  local tool=$1 label=$2 limit=$3; shift 3
  timeout "$limit" compute-sanitizer --tool "$tool" --error-exitcode 99 --report-api-errors no --print-limit 20 \
      python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > "$OUT/${tool}_${label}.log" 2>&1
  local rc=$?
  echo "$tool $label rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/${tool}_${label}.log" | tail -1) : $(grep -E 'passed|failed|error' "$OUT/${tool}_${label}.log" | tail -1)" | tee -a "$OUT/summary.txt"
}
run memcheck pack_crops_eval 300 tests/test_gpu_pack.py tests/test_gpu_crops.py tests/test_gpu_eval.py tests/test_gpu_wb_jitter.py tests/test_realdata_rows.py -k "$SEL"
run memcheck noise 400 tests/test_gpu_noise.py -k "replay_bit_exact_vs_reference or philox_kernel_uses or shard_independence or row_noise or reference_signatures or tail_refinement"
run memcheck unet 400 tests/test_gpu_unet.py -k "$SEL"
run memcheck train 400 tests/test_gpu_train.py -k "act_backward or adam or conv_transpose_backward or head_backward or l1_loss or maxpool_backward or wgrad_nhwc"
run memcheck round2_kernels 500 tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_train_resunet.py -k "fused_first and not 1424 and not 512 or tf32 and not under_any_init or pipeline_equals or resunet_backward"
run racecheck first_layer 400 tests/test_gpu_unet.py -k "fused_first and not 1424 and not 512"
run racecheck noise_eval 400 tests/test_gpu_noise.py tests/test_gpu_eval.py -k "philox_kernel_uses or shard_independence or psnr_ssim or identical"
run racecheck unet_train 400 tests/test_gpu_unet.py tests/test_gpu_train.py -k "conv3x3_layer or fused_pool or x_shift or head_backward or maxpool_backward or l1_loss or act_backward"
cat "$OUT/summary.txt"
