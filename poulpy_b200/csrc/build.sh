#!/usr/bin/env bash
# Builds poulpy_b200/libpoulpy_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# A failed compile aborts the build: every background nvcc is waited for by PID, and the object is removed before
# its compile starts so a failure can never leave a stale .o for the link step.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libpoulpy_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
mkdir -p ../_build
pids=()
names=()
for f in *.cu; do
  o=../_build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/poulpy_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    rm -f "$o"
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -c "$f" -o "$o" &
    pids+=($!)
    names+=("$f")
  fi
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then
    echo "build.sh: nvcc failed on ${names[$i]}" >&2
    rm -f "../_build/${names[$i]%.cu}.o"
    fail=1
  fi
done
if [ "$fail" -ne 0 ]; then
  rm -f "$OUT"
  exit 1
fi
# every source must have an object (a source removed from the tree must not leave its object in the link either)
for o in ../_build/*.o; do
  [ -f "$(basename "${o%.o}").cu" ] || rm -f "$o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT ../_build/*.o -lcudart
echo "built $OUT"
