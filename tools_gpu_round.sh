#!/bin/bash
# One GPU-box round: parity tests, bench, ncu launch list, ncu full captures of the top kernels.
# usage: bash tools_gpu_round.sh <tag> [tests] [ncu]
TAG=${1:-r01}
mkdir -p gpurun_out
if [[ "$2" != "notests" ]]; then
for f in test_gpu_gemm test_gpu_elementwise test_gpu_attention test_gpu_encoder; do
  echo "=== $f"; timeout 700 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/${TAG}_$f.log
done
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "clocks", d["clocks"])
for k,v in d["kernels"].items(): print(f"  {k:24s} {v['launches_per_step']:4d} {v['ms_per_step']:8.3f} ms  {v.get('achieved','')} {v.get('unit','')} {v.get('frac','')}")
print("cpu", d["cpu_baseline"])
PY
if [[ "$3" == "ncu" ]]; then
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 300 -c 12 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn_tc -s 20 -c 2 -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"depthwise|layernorm|subsample_first" -s 30 -c 6 -o gpurun_out/${TAG}_mem $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu mem rc=$?"
ls -la gpurun_out | tail -12
fi
