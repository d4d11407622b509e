python tools/bench_ln.py 2>&1 | head -1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r03d_ln4.json 2> gpurun_out/r03d_ln4.err
python tools/show_bench.py gpurun_out/r03d_ln4.json 2>/dev/null | grep "value\|norm"
