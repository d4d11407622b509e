#!/bin/bash
# same-box A/B of the attention kernels alone: lib/libcfb_prev.so against lib/libcfb.so
L=conformer-nemo_b200/lib
cp $L/libcfb.so $L/libcfb_new.so
for rep in 1 2; do
for which in prev new; do
  cp $L/libcfb_$which.so $L/libcfb.so
  for shape in "16 500 8 64" "32 500 8 64" "128 100 4 64" "1 7500 8 64"; do
    echo -n "$which "; timeout 200 python tools/bench_attn.py $shape 2>/dev/null | head -1
  done
done; done
cp $L/libcfb_new.so $L/libcfb.so
