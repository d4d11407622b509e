#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_packed.py tests/test_gpu_encoder.py -x -q 2>&1 | tail -6 > gpurun_out/r6f_tests.log; cat gpurun_out/r6f_tests.log
python tools/bench_ln.py 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r6f_cfg2.json 2> gpurun_out/r6f.err
python tools/show_bench.py gpurun_out/r6f_cfg2.json | grep -E "value|norm"
