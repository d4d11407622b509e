CFB_ATTN_TRACE=1 python tools/bench_attn.py 2>&1 | tee gpurun_out/r02r_attn_trace.log
python tools/bench_attn.py 2>&1 | head -1 | tee -a gpurun_out/r02r_attn_trace.log
