#!/bin/sh
# Build libcosmoslik_b200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${CSL_OUT:-"$HERE/../libcosmoslik_b200.so"}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  ${CSL_PTXAS_V:+-Xptxas -v} ${CSL_EXTRA} \
  -o "$OUT" "$HERE/api.cu" "$HERE/prepare.cu" "$HERE/binning.cu" "$HERE/chi2.cu" "$HERE/covchi2.cu" "$HERE/samplers.cu"
echo "built $OUT"
