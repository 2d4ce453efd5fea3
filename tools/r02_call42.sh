#!/bin/bash
# round 2, GPU call 42: the warp-specialised SpMM variants ten times over (a producer that owned no stage in 8 consecutive short items
# deadlocked against the scheduler: fixed), then the full GPU suite
set -x
fails=0
for i in 1 2 3 4 5 6 7 8 9 10; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "segmented and (216 or 232 or 316 or 332 or 432 or 516 or 532 or 8t)" > gpurun_out/r02_pytest42_$i.log 2>&1 || fails=$((fails+1))
  tail -1 gpurun_out/r02_pytest42_$i.log
done
echo "failed rounds: $fails"
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest42.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest42.log
