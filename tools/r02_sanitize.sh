#!/bin/bash
# usage: r02_sanitize.sh memcheck|racecheck   (one tool per gpurun call, B200_PROFILING.md)
set -x
TOOL=$1
timeout 300 python tools/smoke.py > gpurun_out/r02_smoke_plain.log 2>&1 || { echo "plain smoke failed"; tail -5 gpurun_out/r02_smoke_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/smoke.py > gpurun_out/r02_sanitizer_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?" >> gpurun_out/r02_sanitizer_$TOOL.log
tail -12 gpurun_out/r02_sanitizer_$TOOL.log | cut -c1-300
