#!/bin/bash
# LayerNorm variants: micro-benchmark + in-situ (cfg2 bench kernels table)
TAG=${1:-r5c}
mkdir -p gpurun_out
for v in "1 2" "2 2" "2 3" "2 4"; do set -- $v; echo "LN_V=$1 BPS=$2"; CFB_LN_V=$1 CFB_LN_BPS=$2 python tools/bench_ln.py 2>&1 | grep layernorm; done
python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_packed.py -x -q -p no:cacheprovider 2>&1 | tail -3
for v in 1 2; do
CFB_LN_V=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${TAG}_cfg2_ln$v.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_cfg2_ln$v.json"))
print("LN_V=$v", round(d["value"]), round(d["ms_per_step"],3), {k:(v["ms_per_step"],v.get("frac")) for k,v in d["kernels"].items() if k.startswith("norm")})
PY
done
