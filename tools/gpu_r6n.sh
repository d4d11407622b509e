#!/bin/bash
# final-state check of the round: GPU test suite, smoke, the default bench line and the reference arm
TAG=${1:-r6n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
(time timeout 800 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err) 2>&1 | grep real
python tools/show_bench.py gpurun_out/${TAG}_bench_cfg2.json | head -3
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_cfg2.json').read().strip().splitlines()[-1]);print(json.dumps(d['pipelines'])[:700]);print('strong', d['strong']['value'], {k:v['value'] for k,v in d['strong']['emulated_on_one_gpu'].items()})"
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err) 2>&1 | grep real
cut -c1-600 gpurun_out/${TAG}_reference_arm.json
