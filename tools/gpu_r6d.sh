#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_ln.py -x -q 2>&1 | tail -15 > gpurun_out/r6d_test_gemm_ln.log; cat gpurun_out/r6d_test_gemm_ln.log
CFB_LNC_TRACE=1 timeout 300 python tools/bench_gemm_ln.py > gpurun_out/r6d_bench_gemm_ln.log 2>&1; cat gpurun_out/r6d_bench_gemm_ln.log
