#!/bin/bash
# (Historical: this is the script as it ran at the start of round 2, against the round-1 defaults.  The variants it timed are defaults
# now, tests/test_gpu_variants.py became tests/test_gpu_variants.py and runs in the normal -m gpu suite; results:
# profiles/r02_sweep_summary.txt.  To re-run it against today's tree, set every PNNP_* variant switch to 0 for the "default" label.)
# Round-2 first GPU call: correctness of the opt-in kernel variants written without a GPU at the end of round 1, then their timing.
#   gpurun --timeout 1500 -- 'bash tools/r02_sweep.sh'      (about 13 minutes of box time)
# Everything lands in gpurun_out/r02_sweep/.  Order: cheap correctness first, so a hang / failure is seen before time is spent.
set -u
OUT=gpurun_out/r02_sweep
mkdir -p "$OUT"
# hand-shake micro-benchmarks (seconds): what one trip through the producer/consumer mbarrier ring costs, per signalling scheme
if [[ -x tools/_bin/ubench_pipeline ]]; then timeout 120 tools/_bin/ubench_pipeline > "$OUT/ubench_pipeline.txt" 2>&1; echo "ubench rc=$?" | tee -a "$OUT/summary.txt"; fi
export PNNP_TEST_EXPERIMENTAL=1
timeout 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_wb_jitter.py tests/test_gpu_preprocess_route.py -q > "$OUT/pytest_experimental.log" 2>&1
echo "experimental tests rc=$?" | tee -a "$OUT/summary.txt"
unset PNNP_TEST_EXPERIMENTAL
run() {  # label, env assignments...
  local label=$1; shift
  ( export "$@"; timeout 200 python tools/profile_unet.py > "$OUT/layers_$label.txt" 2>&1
    timeout 200 python bench.py --workload unet_sony --steps 50 --warmup 5 --no-cpu-baseline > "$OUT/bench_unet_$label.json" 2> "$OUT/bench_unet_$label.err" )
  echo "$label: $(tail -1 "$OUT/layers_$label.txt")" | tee -a "$OUT/summary.txt"
}
run default PNNP_NOOP=1
run super1 PNNP_CONV_SUPER=1
run super2 PNNP_CONV_SUPER=2
run convt PNNP_CONVT_FAST=1
run inv2 PNNP_IN_V2=1
run pdl PNNP_CONV_PDL=1
run x2 PNNP_CONV_F32X2=1
run all1 PNNP_CONV_SUPER=1 PNNP_CONVT_FAST=1 PNNP_IN_V2=1 PNNP_CONV_PDL=1 PNNP_CONV_F32X2=1
run all2 PNNP_CONV_SUPER=2 PNNP_CONVT_FAST=1 PNNP_IN_V2=1 PNNP_CONV_PDL=1 PNNP_CONV_F32X2=1
( export PNNP_COPY_V2=1; timeout 300 python bench.py --workload train_step --steps 30 --warmup 5 --no-cpu-baseline \
    > "$OUT/bench_train_copyv2.json" 2> "$OUT/bench_train_copyv2.err" )
( export PNNP_ACTBWD_V2=1; timeout 300 python bench.py --workload train_step --steps 30 --warmup 5 --no-cpu-baseline \
    > "$OUT/bench_train_actbwdv2.json" 2> "$OUT/bench_train_actbwdv2.err" )
( export PNNP_WGRAD_V2=1; timeout 300 python bench.py --workload train_step --steps 30 --warmup 5 --no-cpu-baseline \
    > "$OUT/bench_train_wgradv2.json" 2> "$OUT/bench_train_wgradv2.err" )
for v in 0 1 2; do
  ( export PNNP_CONV_SUPER=$v PNNP_CONVT_FAST=$((v>0)); timeout 300 python bench.py --workload train_step --steps 30 --warmup 5 --no-cpu-baseline \
      > "$OUT/bench_train_super$v.json" 2> "$OUT/bench_train_super$v.err" )
done
for v in 0 1; do
  ( export PNNP_SSIM_V2=$v; timeout 200 python bench.py --workload sony_evaltest --steps 30 --warmup 5 --no-cpu-baseline \
      > "$OUT/bench_evaltest_ssim$v.json" 2> /dev/null )
  echo "sony_evaltest PNNP_SSIM_V2=$v: $(python -c "import json; d=json.load(open('$OUT/bench_evaltest_ssim$v.json')); print(d['ms_per_step'], 'ms/frame')" 2>/dev/null)" | tee -a "$OUT/summary.txt"
done
# end-to-end synthesis (pinned host in / out): chunk size x stream count of HostSynthPipeline (defaults 8 x 3)
for cfg in "8 3" "4 3" "2 3" "2 4" "4 4" "16 3"; do
  set -- $cfg
  ( export PNNP_E2E_CHUNK=$1 PNNP_E2E_STREAMS=$2; timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
      > "$OUT/bench_synth_chunk$1_streams$2.json" 2> /dev/null )
  echo "synth64 e2e chunk=$1 streams=$2: $(python -c "import json,sys; print(json.load(open('$OUT/bench_synth_chunk$1_streams$2.json'))['e2e']['value'])" 2>/dev/null)" | tee -a "$OUT/summary.txt"
done
for f in "$OUT"/bench_train_*.json; do echo "$(basename "$f"): $(python -c "import json; print(json.load(open('$f'))['ms_per_step'], 'ms/step')" 2>/dev/null)" | tee -a "$OUT/summary.txt"; done
( export PNNP_E2E_ZERO_COPY=1; timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > "$OUT/bench_synth_zero_copy.json" 2> /dev/null )
echo "synth64 e2e zero-copy: $(python -c "import json; print(json.load(open('$OUT/bench_synth_zero_copy.json'))['e2e']['value'])" 2>/dev/null)" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
