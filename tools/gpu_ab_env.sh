#!/bin/bash
# same-box A/B of one environment switch: bash tools/gpu_ab_env.sh VAR A B   (cfg2 step + cfg3 strong shares, two passes)
V=$1; A=$2; B=$3
for rep in 1 2; do
for val in $A $B; do
  env $V=$val timeout 400 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-pipelines 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['strong']
print('$V=$val', 'cfg2 ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), '| strong 1gpu ms', round(s['ms_per_step'],3), 'shares8 ms', round(s['emulated_on_one_gpu']['8']['ms_per_step_slowest_rank'],3), 'shares4', round(s['emulated_on_one_gpu']['4']['ms_per_step_slowest_rank'],3), 'shares2', round(s['emulated_on_one_gpu']['2']['ms_per_step_slowest_rank'],3), '| clocks', d['clocks']['sm_mhz'])"
done; done
