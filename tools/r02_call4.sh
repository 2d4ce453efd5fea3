#!/bin/bash
# round 2, GPU call 4: ncu capture of the segmented SpMM, R-MAT / C4 bench lines, explained-query times, a quick full bench line
set -x
CMD="python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants seg=8"
$CMD > gpurun_out/r02_ncu2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cspmm_seg_kernel" -s 2 -c 1 -o gpurun_out/r02_cspmm_seg $CMD > gpurun_out/r02_ncu2_full.log 2>&1
tail -2 gpurun_out/r02_ncu2_full.log
timeout 600 python tools/explain_query.py --nodes 1000000 --edges 20000000 --graph rmat --communities 500 --queries 3 --device-inputs --profile > gpurun_out/r02_explain_rmat2.txt 2> gpurun_out/r02_explain_rmat2.err
cat gpurun_out/r02_explain_rmat2.txt | cut -c1-420
timeout 900 python bench.py --workload c3_rmat --coalitions 1024 --no-query-leg > gpurun_out/r02_bench_c3_rmat.json 2> gpurun_out/r02_bench_c3_rmat.err; echo "rc=$?" >> gpurun_out/r02_bench_c3_rmat.err
head -c 1200 gpurun_out/r02_bench_c3_rmat.json; echo
timeout 1200 python bench.py --workload c4 --coalitions 128 --cpu-coalitions 1 --no-query-leg > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; echo "rc=$?" >> gpurun_out/r02_bench_c4.err
head -c 1200 gpurun_out/r02_bench_c4.json; echo; tail -3 gpurun_out/r02_bench_c4.err | cut -c1-300
timeout 900 python bench.py --coalitions 1024 > gpurun_out/r02_bench_c3_c.json 2> gpurun_out/r02_bench_c3_c.err; echo "rc=$?" >> gpurun_out/r02_bench_c3_c.err
tail -2 gpurun_out/r02_bench_c3_c.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_c.json')); print(d['value'], d['roofline']['frac']); print(json.dumps(d.get('s_per_explained_query'))[:2500])"
