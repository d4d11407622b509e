#!/bin/bash
# GPU-box round for the transducer greedy decode: parity tests, then the benchmark.  usage: bash tools/gpu_rnnt.sh <tag>
TAG=${1:-r04}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rnnt.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_rnnt.log
timeout 300 python tools/bench_rnnt.py > gpurun_out/${TAG}_bench_rnnt.json 2> gpurun_out/${TAG}_bench_rnnt.err
echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench_rnnt.err; cat gpurun_out/${TAG}_bench_rnnt.json
