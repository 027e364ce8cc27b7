#!/usr/bin/env bash
# Build libmpb_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
SRC=motion_planning_baselines_b200/csrc
OUT=motion_planning_baselines_b200/libmpb_b200.so
mkdir -p build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall $*"
objs=()
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename "${f%.cu}").o
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -name '*.cuh' -newer "$o" -o -name '*.h' -newer "$o")" ]; then
    $NVCC $FLAGS -Xptxas -v -c "$f" -o "$o" 2> "build/$(basename "${f%.cu}").ptxas.log" &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do [ -n "$p" ] && { wait "$p" || fail=1; }; done
if [ "$fail" -ne 0 ]; then grep -h -B2 -A2 "error" build/*.ptxas.log >&2 || true; echo "build failed" >&2; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${objs[@]}" -lcudart
echo "built $OUT"
