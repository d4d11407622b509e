#!/bin/bash
# Round-2 check of the final state: whole GPU suite, smoke, default bench line, reference arm (short)
TAG=${1:-r5g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_cfg2.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg2.json"))
print({k:d[k] for k in ("value","ms_per_step","eager_ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"])
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["sample"])
s=d.get("strong")
print("strong", {k:s[k] for k in ("value","ms_per_step","sub_batches_rank0","launches_per_step_rank0","rank0")}, s["e2e"]["value"])
for w,v in (s.get("emulated_on_one_gpu") or {}).items(): print(" sim", w, round(v["value"]), round(v["speedup_vs_1gpu"],2), v["ms_per_rank"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
echo "reference rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_reference_arm.json')); print(d['value'], d['steps'], d['warmup'], d['cpu_baseline']['kind'], d['cpu_baseline']['how']); print(d.get('cfg1'))"
