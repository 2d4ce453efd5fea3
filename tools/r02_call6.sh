#!/bin/bash
# round 2, GPU call 6: shared-memory ring variant of the segmented SpMM (parity, then timing at C3 / R-MAT)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path" > gpurun_out/r02_pytest6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
tail -4 gpurun_out/r02_pytest6.log
V="seg=8;seg=116;seg=124;seg=16"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var6_c3.jsonl 2> gpurun_out/r02_var6_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 64 --check --variants "$V" > gpurun_out/r02_var6_rmat.jsonl 2> gpurun_out/r02_var6_rmat.err
cat gpurun_out/r02_var6_c3.jsonl gpurun_out/r02_var6_rmat.jsonl; tail -3 gpurun_out/r02_var6_c3.err
