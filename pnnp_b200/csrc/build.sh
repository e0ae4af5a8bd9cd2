#!/bin/bash
# Builds libpnnp_b200.so (sm_100a) in-tree.  Usage: pnnp_b200/csrc/build.sh
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libpnnp_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart shared
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "$HERE/_obj"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/_obj/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find "$HERE" "$HERE/../../include" -name '*.h' -newer "$o" -o -name '*.cuh' -newer "$o")" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; rm -f "$o"; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do if [[ -n "$p" ]]; then wait "$p" || rc=1; fi; done
if [[ $rc -ne 0 ]]; then echo "build.sh: compilation FAILED" >&2; rm -f "$OUT"; exit 1; fi
"$NVCC" -shared -cudart shared -o "$OUT" "$HERE"/_obj/*.o
echo "built $OUT"
# developer tools (micro-benchmarks run under gpurun; not part of the library): built when stale, never fatal
TOOLS="$HERE/../../tools"
if [[ -f "$TOOLS/ubench_pipeline.cu" && ( ! -x "$TOOLS/_bin/ubench_pipeline" || "$TOOLS/ubench_pipeline.cu" -nt "$TOOLS/_bin/ubench_pipeline" ) ]]; then
  mkdir -p "$TOOLS/_bin"
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o "$TOOLS/_bin/ubench_pipeline" "$TOOLS/ubench_pipeline.cu" > /dev/null 2>&1 || echo "build.sh: tools/ubench_pipeline not built (ignored)"
fi
