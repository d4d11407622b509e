#!/bin/bash
# One GPU-box round: parity tests, bench, and (optionally) the ncu launch list + full captures of the top kernels.
# usage: bash tools_gpu_round.sh <tag> [tests|notests] [ncu]
TAG=${1:-r01}
mkdir -p gpurun_out
if [[ "$2" != "notests" ]]; then
for f in test_gpu_conv_tail test_gpu_gemm test_gpu_gemm_ln test_gpu_elementwise test_gpu_attention test_gpu_encoder test_ctc_head; do
  echo "=== $f"; timeout 700 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/${TAG}_$f.log
done
fi
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [[ "$3" == "ncu" ]]; then
CMD="python bench.py --ncu"
LPS=245   # launches per cfg2 step (bench.py --ncu prints it)
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 300 -c 14 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn_tc -s 20 -c 2 -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"depthwise|layernorm|subsample_first" -s 30 -c 8 -o gpurun_out/${TAG}_mem $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu mem rc=$?"
ls -la gpurun_out | tail -12
fi
