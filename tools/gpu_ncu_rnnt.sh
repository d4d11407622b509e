TAG=${1:-r04x}
timeout 300 python tools/bench_rnnt.py > gpurun_out/${TAG}_bench_rnnt.json 2> gpurun_out/${TAG}_bench_rnnt.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_rnnt_launches.csv \
  python tools/bench_rnnt.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/${TAG}_rnnt_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rnnt_greedy -s 13 -c 1 -o gpurun_out/${TAG}_rnnt \
  python tools/bench_rnnt.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/${TAG}_rnnt_ncu2.log 2>&1
echo "ncu full rc=$?"
