#!/bin/bash
# Round-2 evidence pass: ncu launch list of one cfg2 step, --set full captures of the top kernels (shipped build), the
# same for an 8-GPU share of cfg3 (packed path), env probe under ncu, canary tests.
TAG=${1:-r5f}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_canary.py -x -q -p no:cacheprovider 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum python -c "
import os
for k,v in sorted(os.environ.items()):
    if any(s in k.upper() for s in ('NSIGHT','INJECT','NCU','NV_','CUDA','LD_PRELOAD','PROFIL')): print('ENV',k,'=',v[:160])
" 2>&1 | grep ENV | head -20
CMD="python bench.py --ncu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
LPS=$(python -c "import json,sys; print([json.loads(l)['launches_per_step'] for l in open('gpurun_out/${TAG}_plain.log') if l.startswith('{')][-1])")
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
export CFB_MICROBATCH=0
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 150 -c 24 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn_tc -s 10 -c 2 -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"dw_pw|layernorm|conv0_im2col" -s 15 -c 10 -o gpurun_out/${TAG}_mem $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu mem rc=$?"
unset CFB_MICROBATCH
# the packed path on an 8-GPU share of cfg3: launch list
CMD3="python bench.py --ncu --workload cfg3 --share 0/8"
$CMD3 > gpurun_out/${TAG}_plain3.log 2>&1
LPS3=$(python -c "import json,sys; print([json.loads(l)['launches_per_step'] for l in open('gpurun_out/${TAG}_plain3.log') if l.startswith('{')][-1])")
echo "launches per step (cfg3 share 0/8): $LPS3"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*LPS3)) -c $LPS3 --csv --log-file gpurun_out/${TAG}_launches_cfg3_share8.csv $CMD3 > gpurun_out/${TAG}_ncu5.log 2>&1
echo "ncu launches cfg3 rc=$?"
ls -la gpurun_out | grep ${TAG}
