#!/bin/bash
# Build libpocketnerf.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off --expt-relaxed-constexpr ${PN_NVCC_EXTRA:-}"
# PN_NVCC_EXTRA / PN_BUILD_DIR / PN_LIB_NAME: a second build with other -D tuning constants next to the default one
BUILD=${PN_BUILD_DIR:-build}
LIBNAME=${PN_LIB_NAME:-libpocketnerf.so}
SRCS="api.cu hash_encode.cu composite.cu sample.cu rays.cu mlp_fp32.cu"
for extra in mlp_tc.cu fused.cu optim.cu dataio.cu; do [ -f "$extra" ] && SRCS="$SRCS $extra"; done
mkdir -p "$BUILD"
objs=""
pids=""
for s in $SRCS; do
  o="$BUILD/${s%.cu}.o"
  objs="$objs $o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o")" ] || [ ../../include/pocketnerf.h -nt "$o" ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$s" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait "$p"; done
$NVCC -shared -o "$LIBNAME" $objs -lcudart
echo "built $(pwd)/$LIBNAME"
