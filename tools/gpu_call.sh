timeout 600 python -m pytest tests/test_frontend.py -x -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/r02y_test_frontend.log
python tools/bench_frontend.py | tee gpurun_out/r02z_bench_frontend.json
