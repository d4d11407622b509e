timeout 600 python -m pytest tests/test_gpu_conv_tail.py -x -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -15
CFB_TAIL_TRACE=1 python tools/bench_conv_tail.py 2>&1 | tee gpurun_out/r02f_trace.log
python tools/bench_conv_tail.py 2>&1 | tee gpurun_out/r02f_ubench.log
python tools/bench_conv_tail.py 256 100 256 2>&1 | tee -a gpurun_out/r02f_ubench.log
python tools/bench_conv_tail.py 1 7500 512 2>&1 | tee -a gpurun_out/r02f_ubench.log
