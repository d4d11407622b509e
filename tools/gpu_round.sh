#!/bin/bash
# One GPU-box round: parity tests (the driver's own command), bench, and (optionally) the ncu launch list + full
# captures of the top kernels.  usage: bash tools/gpu_round.sh <tag> [tests|notests] [ncu]
TAG=${1:-r01}
mkdir -p gpurun_out
if [[ "$2" != "notests" ]]; then
  timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.log
fi
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [[ "$3" == "ncu" ]]; then
CMD="python bench.py --ncu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
LPS=$(python -c "import json,sys; print([json.loads(l)['launches_per_step'] for l in open('gpurun_out/${TAG}_plain.log') if l.startswith('{')][-1])")
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
# the --set full captures run the forward as ONE batch (CFB_MICROBATCH=0), like bench.py's per-kernel timing pass, so
# per-launch figures (DRAM bytes, durations) refer to the same launches as roofline.achieved
export CFB_MICROBATCH=0
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 150 -c 24 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn_tc -s 10 -c 2 -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"dw_pw|layernorm|conv0_im2col" -s 15 -c 10 -o gpurun_out/${TAG}_mem $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu mem rc=$?"
ls -la gpurun_out | tail -12
fi
