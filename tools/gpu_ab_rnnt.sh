#!/bin/bash
# same-box A/B of the transducer decode kernels: lib/libcfb_prev.so against lib/libcfb.so (cluster kernel and, with
# CFB_RNNT_CLUSTER=0, the row-partitioned one)
L=conformer-nemo_b200/lib
cp $L/libcfb.so $L/libcfb_new.so
for rep in 1 2; do
for which in prev new; do
  cp $L/libcfb_$which.so $L/libcfb.so
  for c in 1 0; do
    echo -n "$which cluster=$c "; CFB_RNNT_CLUSTER=$c timeout 200 python tools/bench_rnnt.py --cpu-sample 0 --steps 10 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in d if k in ('ms_per_step','ms_per_batch','value','iterations','symbols')})"
  done
done; done
cp $L/libcfb_new.so $L/libcfb.so
