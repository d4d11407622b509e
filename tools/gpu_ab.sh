#!/bin/bash
# same-box A/B of two builds of the library: lib/libcfb_prev.so against lib/libcfb.so (cfg2 step + cfg3 strong shares)
L=conformer-nemo_b200/lib
cp $L/libcfb.so $L/libcfb_new.so
for rep in 1 2; do
for which in prev new; do
  cp $L/libcfb_$which.so $L/libcfb.so
  timeout 400 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-pipelines 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['strong']
print('$which', 'cfg2 ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), '| gemm frac', d['roofline']['frac'], '| strong 1gpu ms', round(s['ms_per_step'],3), 'shares8 ms', round(s['emulated_on_one_gpu']['8']['ms_per_step_slowest_rank'],3), 'shares4', round(s['emulated_on_one_gpu']['4']['ms_per_step_slowest_rank'],3), '| clocks', d['clocks']['sm_mhz'])"
done; done
cp $L/libcfb_new.so $L/libcfb.so
