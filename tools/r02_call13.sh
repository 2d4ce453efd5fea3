#!/bin/bash
# round 2, GPU call 13: warp-specialised TMA bulk-copy SpMM (compact_bulk.cu; seg = 216 / 232 / 248): parity, then timing against seg = 8
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented" > gpurun_out/r02_pytest13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest13.log
tail -6 gpurun_out/r02_pytest13.log
V="seg=8;seg=248;seg=232;seg=216"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var13_c3.jsonl 2> gpurun_out/r02_var13_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg=8;seg=248" > gpurun_out/r02_var13_rmat.jsonl 2> gpurun_out/r02_var13_rmat.err
cat gpurun_out/r02_var13_c3.jsonl gpurun_out/r02_var13_rmat.jsonl; tail -3 gpurun_out/r02_var13_c3.err
