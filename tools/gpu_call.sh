timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -12 | tee gpurun_out/r02u_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
python tools/show_bench.py gpurun_out/r02u_bench.json | head -4
python bench.py --steps 20 --warmup 3 --workload cfg4 --no-cpu-baseline > gpurun_out/r02u_bench_cfg4.json 2> gpurun_out/r02u_bench_cfg4.err
python tools/show_bench.py gpurun_out/r02u_bench_cfg4.json | head -3
CFB_ATTN_PERSIST=0 python bench.py --steps 20 --warmup 3 --workload cfg4 --no-cpu-baseline > gpurun_out/r02u_bench_cfg4_np.json 2> gpurun_out/r02u_bench_cfg4_np.err
python tools/show_bench.py gpurun_out/r02u_bench_cfg4_np.json | head -3
