#!/bin/bash
# Round-end GPU pass: the driver's own test command, the headline bench, the transducer decode bench, and ncu evidence
# (launch list + one --set full capture) for the decode kernel.  usage: bash tools/gpu_round4.sh <tag>
TAG=${1:-r04}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_cfg2.err; python tools/show_bench.py gpurun_out/${TAG}_bench_cfg2.json 2>/dev/null | head -20
timeout 300 python tools/bench_rnnt.py > gpurun_out/${TAG}_bench_rnnt.json 2> gpurun_out/${TAG}_bench_rnnt.err
echo "rnnt bench rc=$?"; cat gpurun_out/${TAG}_bench_rnnt.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_rnnt_launches.csv \
  python tools/bench_rnnt.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/${TAG}_rnnt_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rnnt_greedy -s 13 -c 1 -o gpurun_out/${TAG}_rnnt \
  python tools/bench_rnnt.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/${TAG}_rnnt_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
