#!/bin/bash
mkdir -p gpurun_out
CFB_FUSED_LN=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r6e_ln0.json 2> gpurun_out/r6e_ln0.err
CFB_FUSED_LN=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r6e_ln1.json 2> gpurun_out/r6e_ln1.err
python - <<'P'
import json
for f in ("gpurun_out/r6e_ln0.json","gpurun_out/r6e_ln1.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    s=d.get("strong",{})
    print(f, "cfg2 ms", round(d["ms_per_step"],3), "| strong 1gpu ms", round(s.get("ms_per_step",0),3), "| emulated:", json.dumps(s.get("emulated_on_one_gpu"))[:600])
P
