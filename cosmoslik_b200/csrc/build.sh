#!/bin/sh
# Build libcosmoslik_b200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
# Every translation unit is compiled in its own nvcc process (in parallel), then linked.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${CSL_OUT:-"$HERE/../libcosmoslik_b200.so"}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OBJ=${CSL_OBJDIR:-"$HERE/../../build/obj"}
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${CSL_PTXAS_V:+-Xptxas -v} ${CSL_EXTRA}"
UNITS="api prepare binning chi2 covchi2 samplers mh"
pids=""
for u in $UNITS; do
  # rebuild a unit only when it, a header or the flags changed (stamp = flags)
  stamp="$OBJ/$u.flags"
  if [ -f "$OBJ/$u.o" ] && [ -f "$stamp" ] && [ "$(cat "$stamp")" = "$FLAGS" ] && \
     [ -z "$(find "$HERE" "$HERE/../../include" -newer "$OBJ/$u.o" \( -name '*.cuh' -o -name '*.h' -o -name "$u.cu" \) 2>/dev/null)" ]; then
    continue
  fi
  ( $NVCC $FLAGS -c -o "$OBJ/$u.o" "$HERE/$u.cu" > "$OBJ/$u.log" 2>&1 && echo "$FLAGS" > "$stamp" ) &
  pids="$pids $!:$u"
done
fail=0
for pu in $pids; do
  pid=${pu%%:*}; u=${pu##*:}
  if ! wait "$pid"; then fail=1; echo "---- $u.cu failed"; fi
  cat "$OBJ/$u.log"
done
[ "$fail" = 0 ] || exit 1
objs=""
for u in $UNITS; do objs="$objs $OBJ/$u.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $objs
echo "built $OUT"
