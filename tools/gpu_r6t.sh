#!/bin/bash
# final build with the fp16 position term: GPU suite, bench, launch list + full capture of the attention kernels
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r6t_pytest_gpu.log; cat gpurun_out/r6t_pytest_gpu.log
(time timeout 800 python bench.py > gpurun_out/r6t_bench_cfg2.json 2> gpurun_out/r6t_bench.err) 2>&1 | grep real
python tools/show_bench.py gpurun_out/r6t_bench_cfg2.json 2>/dev/null | head -22
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/r6t_bench_cfg5.json 2>/dev/null; python tools/show_bench.py gpurun_out/r6t_bench_cfg5.json 2>/dev/null | grep -E "value|attention"
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/r6t_bench_cfg4.json 2>/dev/null; python tools/show_bench.py gpurun_out/r6t_bench_cfg4.json 2>/dev/null | grep -E "value|attention"
# ncu: launch list of one cfg2 step, then a full capture of two attention launches (persistent kernel) and of the per-item kernel on cfg5
python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6t_ncu_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1374 -c 458 --csv --log-file gpurun_out/r6t_launches.csv python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6t_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn -s 51 -c 2 -o gpurun_out/r6t_attn python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6t_ncu_attn.log 2>&1; echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn -s 51 -c 1 -o gpurun_out/r6t_attn_cfg5 python bench.py --ncu --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6t_ncu_attn5.log 2>&1; echo "attn cfg5 capture rc=$?"
ls -la gpurun_out/r6t_*
