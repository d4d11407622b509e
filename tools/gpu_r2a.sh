#!/bin/bash
# Round-2 GPU pass A: packed-path parity, the whole GPU suite, the default bench line (with the strong sub-record).
TAG=${1:-r5a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
timeout 900 python -m pytest tests/test_gpu_packed.py -x -q -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_packed.log
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_packed.py 2>&1 | grep -v "^E  " | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_cfg2.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg2.json"))
print({k:d[k] for k in ("value","ms_per_step","eager_ms_per_step")}, d["e2e"]["value"], d["cpu_baseline"])
s=d.get("strong")
if s:
    print("strong", {k:s[k] for k in ("value","ms_per_step","sub_batches_rank0","launches_per_step_rank0","rank0")}, s["e2e"]["value"])
    for w,v in (s.get("emulated_on_one_gpu") or {}).items(): print(" sim", w, v)
for k,v in d["kernels"].items(): print(" ", k, v)
PY
timeout 600 python bench.py --workload cfg3 --packed off --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_dense.json 2> gpurun_out/${TAG}_bench_cfg3_dense.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_cfg3_dense.json')); print('cfg3 dense', d['value'], d['ms_per_step'], d['e2e']['value'])"
