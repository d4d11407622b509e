#!/bin/bash
# transducer decode + whole pipelines with the final build (cluster decode kernel is the default now)
TAG=${1:-r5n}
mkdir -p gpurun_out
timeout 300 python tools/bench_rnnt.py > gpurun_out/${TAG}_bench_rnnt.json 2> gpurun_out/${TAG}_rnnt.err; cat gpurun_out/${TAG}_bench_rnnt.json | cut -c1-700
CFB_RNNT_CLUSTER=0 timeout 300 python tools/bench_rnnt.py --cpu-sample 0 > gpurun_out/${TAG}_bench_rnnt_rowpart.json 2>> gpurun_out/${TAG}_rnnt.err; cat gpurun_out/${TAG}_bench_rnnt_rowpart.json | cut -c1-400
timeout 600 python tools/bench_pipeline.py > gpurun_out/${TAG}_bench_pipeline.json 2> gpurun_out/${TAG}_pipe.err; cat gpurun_out/${TAG}_bench_pipeline.json | cut -c1-1500
