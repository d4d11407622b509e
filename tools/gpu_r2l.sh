#!/bin/bash
TAG=${1:-r5l}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_packed.py tests/test_gpu_gemm.py tests/test_gpu_canary.py -x -q -p no:cacheprovider 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:round(d[k],3) for k in ("value","ms_per_step","eager_ms_per_step")}, round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
PY
}
for bn in 1 0; do for sh in 0/8 3/8 0/4; do
  CFB_GEMM_SMALL_BN=$bn timeout 300 python bench.py --workload cfg3 --share $sh --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg3_bn${bn}_${sh/\//of}.json 2> gpurun_out/${TAG}_err.log || tail -3 gpurun_out/${TAG}_err.log
  show gpurun_out/${TAG}_cfg3_bn${bn}_${sh/\//of}.json
done; done
for bn in 1 0; do CFB_GEMM_SMALL_BN=$bn python bench.py --workload cfg1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg1_bn${bn}.json 2>/dev/null; show gpurun_out/${TAG}_cfg1_bn${bn}.json; done
