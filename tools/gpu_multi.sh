#!/bin/bash
# Multi-GPU pass: the default line (cfg2 weak + cfg3 strong sub-record) and cfg4 (weak) at N GPUs.  usage: gpu_multi.sh N tag
N=${1:-8}; TAG=${2:-r5m}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@"; }
run > gpurun_out/${TAG}_bench_cfg2_${N}gpu.json 2> gpurun_out/${TAG}_cfg2_${N}gpu.err || tail -5 gpurun_out/${TAG}_cfg2_${N}gpu.err
run --workload cfg4 > gpurun_out/${TAG}_bench_cfg4_${N}gpu.json 2> gpurun_out/${TAG}_cfg4_${N}gpu.err || tail -5 gpurun_out/${TAG}_cfg4_${N}gpu.err
python - <<PY
import json
for w in ("cfg2","cfg4"):
    try:
        d=json.load(open(f"gpurun_out/${TAG}_bench_{w}_${N}gpu.json"))
    except Exception as e:
        print(w, "failed", e); continue
    print(w, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["clocks"])
    s=d.get("strong")
    if s: print(" strong", round(s["value"]), round(s["ms_per_step"],3), "e2e", round(s["e2e"]["value"]), s["sub_batches_rank0"], s["rank0"])
PY
