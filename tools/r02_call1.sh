#!/bin/bash
# round 2, GPU call 1: parity suite with the segmented SpMM, then the variant sweep at C3 / R-MAT
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
V="seg=0;seg=4;seg=4,seg_occ=10;seg=6;seg=6,seg_occ=6;seg=8;seg=8,seg_occ=8"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var_c3.jsonl 2> gpurun_out/r02_var_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 64 --check --variants "$V" > gpurun_out/r02_var_rmat.jsonl 2> gpurun_out/r02_var_rmat.err
tail -3 gpurun_out/r02_pytest1.log; cat gpurun_out/r02_var_c3.jsonl gpurun_out/r02_var_rmat.jsonl
