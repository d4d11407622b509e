#!/bin/bash
# single-GPU records of the other BASELINE configs (cfg1, cfg4, cfg5) with the final build
TAG=${1:-r5k}
mkdir -p gpurun_out
for w in cfg1 cfg4 cfg5; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}.json 2> gpurun_out/${TAG}_${w}.err || tail -3 gpurun_out/${TAG}_${w}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${w}.json"))
print("$w", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "eager", round(d["eager_ms_per_step"],3), d["roofline"]["frac"])
PY
done
