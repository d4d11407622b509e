#!/bin/bash
# final-state check after the attention scheduling changes: GPU suite, smoke, default bench line, cfg 4 / cfg 5 lines, reference arm
TAG=${1:-r7a}
bash tools/gpu_r6n.sh $TAG
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${TAG}_bench_cfg5.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_cfg5.json 2>/dev/null | grep -E "value|attention"
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${TAG}_bench_cfg4.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_cfg4.json 2>/dev/null | grep -E "value|attention"
