for wl in cfg3 cfg5 cfg4 tiny; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r03a_bench_$wl.json 2> gpurun_out/r03a_bench_$wl.err
  echo "$wl rc=$?"; python tools/show_bench.py gpurun_out/r03a_bench_$wl.json 2>/dev/null | head -2 | tail -1
done
