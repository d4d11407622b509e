#!/bin/bash
# same-box A/B of the GEMM family alone: lib/libcfb_prev.so against lib/libcfb.so on the cfg 2 and cfg 4 layer shapes
L=conformer-nemo_b200/lib
cp $L/libcfb.so $L/libcfb_new.so
for rep in 1 2; do
for which in prev new; do
  cp $L/libcfb_$which.so $L/libcfb.so
  echo "== $which cfg2"; timeout 200 python tools/bench_gemm.py 2>/dev/null | grep -E "us" | grep -v "^ " | cut -c1-110
  echo "== $which cfg4"; timeout 200 python tools/bench_gemm.py cfg4 2>/dev/null | grep -E "us" | grep -v "^ " | cut -c1-110
done; done
cp $L/libcfb_new.so $L/libcfb.so
