#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_packed.py -x -q 2>&1 | tail -8 > gpurun_out/r6c_tests.log; cat gpurun_out/r6c_tests.log
CFB_FUSED_LN=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r6c_cfg2_ln0.json 2> gpurun_out/r6c_ln0.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r6c_cfg2_ln1.json 2> gpurun_out/r6c_ln1.err
python tools/show_bench.py gpurun_out/r6c_cfg2_ln0.json gpurun_out/r6c_cfg2_ln1.json
tail -3 gpurun_out/r6c_ln1.err
