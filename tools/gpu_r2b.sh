#!/bin/bash
# Round-2 GPU pass B: packed two-group micro-batching; per-kernel profile of an 8-way share of cfg3; PDL at small sizes.
TAG=${1:-r5b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py -x -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_packed.log
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:round(d[k],3) for k in ("value","ms_per_step","eager_ms_per_step")}, round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
PY
}
for mb in 1 0; do
  for sh in 0/1 0/8 0/4; do
    CFB_MICROBATCH=$mb timeout 300 python bench.py --workload cfg3 --share $sh --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg3_mb${mb}_${sh/\//of}.json 2> gpurun_out/${TAG}_err.log || tail -3 gpurun_out/${TAG}_err.log
    show gpurun_out/${TAG}_cfg3_mb${mb}_${sh/\//of}.json
  done
done
CFB_PDL=1 timeout 300 python bench.py --workload cfg3 --share 0/8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg3_pdl_0of8.json 2> gpurun_out/${TAG}_err.log
show gpurun_out/${TAG}_cfg3_pdl_0of8.json
CFB_PDL=1 CFB_MICROBATCH=0 timeout 300 python bench.py --workload cfg3 --share 0/8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg3_pdl_mb0_0of8.json 2> gpurun_out/${TAG}_err.log
show gpurun_out/${TAG}_cfg3_pdl_mb0_0of8.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_cfg3_mb1_0of8.json"))
tot=0
for k,v in d["kernels"].items():
    print(" ", k, v["launches_per_step"], v["ms_per_step"], v.get("frac")); tot+=v["ms_per_step"]
print("sum", tot)
PY
