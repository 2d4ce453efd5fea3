#!/bin/bash
# round 2, GPU call 25: full GPU suite; explained R-MAT queries (hub / median / leaf) with the sliced hub rows in the pruned layer 0
set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest25.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest25.log
tail -3 gpurun_out/r02_pytest25.log
timeout 900 python tools/explain_query.py --nodes 1000000 --edges 20000000 --graph rmat --communities 500 --queries 3 --device-inputs --profile > gpurun_out/r02_explain_rmat2.txt 2> gpurun_out/r02_explain_rmat2.err
cut -c1-420 gpurun_out/r02_explain_rmat2.txt; tail -2 gpurun_out/r02_explain_rmat2.err
