#!/bin/bash
mkdir -p gpurun_out
for pm in 0 1; do CFB_ATTN_PERSIST=$pm timeout 300 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_packed.py tests/test_ctc_head.py -x -q -m gpu 2>&1 | tail -2
for pm in 0 1; do echo "persist=$pm"; CFB_ATTN_PERSIST=$pm python tools/bench_attn.py 2>&1 | tail -2; CFB_ATTN_PERSIST=$pm python tools/bench_attn.py 256 100 4 64 2>&1 | tail -2; done
CFB_ATTN_PERSIST=0 python tools/bench_attn.py 1 7500 8 64 2>&1 | tail -1
bash tools/gpu_ab.sh
