#!/bin/bash
# compute-sanitizer over every product kernel on small shapes (tools/sanitize_smoke.py); logs land in gpurun_out/.
# The seven parts of a tool run side by side (separate processes on the one GPU); each is bounded by SAN_TIMEOUT.
TAG=${1:-r5s}
TOOLS=${2:-"memcheck racecheck"}
mkdir -p gpurun_out
python tools/sanitize_smoke.py > gpurun_out/${TAG}_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/${TAG}_plain.log
PARTS="gemm elementwise attention encoder ctc frontend rnnt"
for tool in $TOOLS; do
  for part in $PARTS; do
    ( timeout ${SAN_TIMEOUT:-300} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py $part > gpurun_out/${TAG}_${tool}_${part}.log 2>&1; echo $? > gpurun_out/${TAG}_${tool}_${part}.rc ) &
  done
  wait
  for part in $PARTS; do
    echo "== $tool $part rc=$(cat gpurun_out/${TAG}_${tool}_${part}.rc) stages=$(grep -c 'stage:' gpurun_out/${TAG}_${tool}_${part}.log) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_${tool}_${part}.log | tail -1)"
  done
done
