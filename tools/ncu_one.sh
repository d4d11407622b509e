#!/bin/bash
# usage: bash tools/ncu_one.sh <tag> <kernel-regex> <skip> <count> [workload]
TAG=$1; REGEX=$2; SKIP=${3:-20}; COUNT=${4:-2}; WL=${5:-cfg2}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $WL"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
