#!/bin/bash
# final build of the round: GPU suite, smoke, bench lines, reference arm, then the ncu launch list of one cfg 2 step and
# full captures of the two attention kernels (each after the same command has exited 0 without ncu)
TAG=${1:-r7b}
bash tools/gpu_r7a.sh $TAG
python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1374 -c 458 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn -s 51 -c 2 -o gpurun_out/${TAG}_attn python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_attn.log 2>&1; echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rel_attn -s 51 -c 1 -o gpurun_out/${TAG}_attn_cfg5 python bench.py --ncu --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_attn5.log 2>&1; echo "attn cfg5 capture rc=$?"
ls -la gpurun_out/${TAG}_*
