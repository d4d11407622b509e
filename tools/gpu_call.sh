python tools/bench_write_bw.py 2>&1 | tee gpurun_out/r02o_write_bw.log
