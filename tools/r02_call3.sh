#!/bin/bash
# round 2, GPU call 3: parity suite (per-test timeout), l0 TMA variant, default bench, hub-query host profile, ncu of the SpMM
set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/r02_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -5 gpurun_out/r02_pytest3.log
V="seg=8;seg=8,l0_ws=3;seg=6,seg_occ=6,l0_ws=3"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var3_c3.jsonl 2> gpurun_out/r02_var3_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 64 --check --variants "$V" > gpurun_out/r02_var3_rmat.jsonl 2> gpurun_out/r02_var3_rmat.err
cat gpurun_out/r02_var3_c3.jsonl gpurun_out/r02_var3_rmat.jsonl
timeout 900 python bench.py > gpurun_out/r02_bench_c3_b.json 2> gpurun_out/r02_bench_c3_b.err
echo "bench rc=$?" >> gpurun_out/r02_bench_c3_b.err
tail -3 gpurun_out/r02_bench_c3_b.err
CUDA_LAUNCH_BLOCKING=1 timeout 600 python tools/explain_query.py --nodes 1000000 --edges 20000000 --graph rmat --communities 500 --queries 2 --device-inputs --hostprof > gpurun_out/r02_explain_hostprof.txt 2> gpurun_out/r02_explain_hostprof.err
timeout 600 python tools/explain_query.py --nodes 1000000 --edges 20000000 --graph rmat --communities 500 --queries 3 --device-inputs --profile > gpurun_out/r02_explain_rmat.txt 2> gpurun_out/r02_explain_rmat.err
cat gpurun_out/r02_explain_rmat.txt
# ncu: the same command first without ncu (exit 0), then the launch list and one full capture of the SpMM / layer-0 kernels
CMD="python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants seg=8"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cspmm_seg_kernel|l0_ws_kernel|dense_tc_kernel" -s 6 -c 3 -o gpurun_out/r02_seg_l0_dense $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
