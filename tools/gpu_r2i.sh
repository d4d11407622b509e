#!/bin/bash
for f in 0 64; do for shape in "32 500 8 64" "1 7500 8 64"; do echo "tc dbg=$f shape=$shape"; CFB_ATTN_PERSIST=0 CFB_ATTN_DEBUG=$f timeout 120 python tools/bench_attn.py $shape 2>&1 | tail -2; done; done
