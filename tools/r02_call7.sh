#!/bin/bash
# round 2, GPU call 7: segmented SpMM v4 (one uniform predicated stream loop), occupancy variants
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path" > gpurun_out/r02_pytest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest7.log
tail -3 gpurun_out/r02_pytest7.log
V="seg=8;seg=8,seg_occ=7;seg=6,seg_occ=6;seg=6,seg_occ=7;seg=6;seg=4;seg=0"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var7_c3.jsonl 2> gpurun_out/r02_var7_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 64 --check --variants "$V" > gpurun_out/r02_var7_rmat.jsonl 2> gpurun_out/r02_var7_rmat.err
cat gpurun_out/r02_var7_c3.jsonl gpurun_out/r02_var7_rmat.jsonl; tail -3 gpurun_out/r02_var7_c3.err
