timeout 600 python -m pytest tests/test_frontend.py -x -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r02y_test_frontend.log
