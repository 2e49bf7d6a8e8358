#!/usr/bin/env bash
# Builds poulpy_b200/libpoulpy_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libpoulpy_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
mkdir -p ../_build
for f in *.cu; do
  o=../_build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/poulpy_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -c "$f" -o "$o" &
  fi
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT ../_build/*.o -lcudart
echo "built $OUT"
