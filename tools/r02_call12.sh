#!/bin/bash
# round 2, GPU call 12: source-range passes of the segmented SpMM (parity, then timing), and the half-size C3 graph (L2 experiment)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale" > gpurun_out/r02_pytest12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest12.log
tail -4 gpurun_out/r02_pytest12.log
V="seg=8;seg=8,seg_half=1;seg=6,seg_occ=6,seg_half=1;seg=4,seg_half=1"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var12_c3.jsonl 2> gpurun_out/r02_var12_c3.err
timeout 600 python tools/variants.py --workload c3_half --coalitions 128 --check --variants "seg=8;seg=0" > gpurun_out/r02_var12_c3half.jsonl 2> gpurun_out/r02_var12_c3half.err
cat gpurun_out/r02_var12_c3.jsonl gpurun_out/r02_var12_c3half.jsonl; tail -3 gpurun_out/r02_var12_c3.err
