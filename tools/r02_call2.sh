#!/bin/bash
# round 2, GPU call 2: parity suite, segmented-SpMM variant sweep (C3 / R-MAT), the default bench line, the reference arm
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
V="seg=0;seg=4;seg=6;seg=6,seg_occ=6;seg=8;seg=8,seg_occ=8"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var2_c3.jsonl 2> gpurun_out/r02_var2_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 64 --check --variants "$V" > gpurun_out/r02_var2_rmat.jsonl 2> gpurun_out/r02_var2_rmat.err
timeout 900 python bench.py > gpurun_out/r02_bench_c3_a.json 2> gpurun_out/r02_bench_c3_a.err
echo "bench rc=$?" >> gpurun_out/r02_bench_c3_a.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_a.json 2> gpurun_out/r02_bench_ref_a.err
tail -3 gpurun_out/r02_pytest2.log; cat gpurun_out/r02_var2_c3.jsonl gpurun_out/r02_var2_rmat.jsonl; tail -5 gpurun_out/r02_bench_c3_a.err; cat gpurun_out/r02_bench_c3_a.json | head -c 6000; cat gpurun_out/r02_bench_ref_a.json | head -c 3000
