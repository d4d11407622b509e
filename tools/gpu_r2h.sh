#!/bin/bash
# attention_tc5 (two threads per row): parity + micro-benchmark + in-situ
TAG=${1:-r5h}
mkdir -p gpurun_out
CFB_ATTN_V=5 timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -p no:cacheprovider 2>&1 | tail -8
for shape in "32 500 8 64" "256 100 4 64" "1 7500 8 64" "8 750 8 64"; do
  for v in 0 5; do echo "V=$v shape=$shape"; CFB_ATTN_V=$v timeout 120 python tools/bench_attn.py $shape 2>&1 | tail -2; done
done
CFB_ATTN_V=5 timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_packed.py -x -q -p no:cacheprovider 2>&1 | tail -4
for v in 0 5; do
CFB_ATTN_V=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-sim > gpurun_out/${TAG}_cfg2_v$v.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_cfg2_v$v.json"))
print("V=$v", round(d["value"]), round(d["ms_per_step"],3), d["kernels"]["rel-pos attention"], "strong", round(d["strong"]["value"]), round(d["strong"]["ms_per_step"],3))
PY
done
