#!/bin/sh
# bench every prebuilt library variant under variants_tmp/ (built here with CSL_OUT / CSL_EXTRA)
cp cosmoslik_b200/libcosmoslik_b200.so /tmp/keep.so
for v in variants_tmp/*.so; do
  cp "$v" cosmoslik_b200/libcosmoslik_b200.so
  echo "== $v"
  python bench.py --steps ${STEPS:-50} --no-cpu-baseline $BENCH_ARGS 2>&1 | python tools/brief.py
done
cp /tmp/keep.so cosmoslik_b200/libcosmoslik_b200.so
