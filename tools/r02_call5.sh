#!/bin/bash
# round 2, GPU call 5: parity suite after the dense_tc epilogue change, deeper gather queues, R-MAT parity values
set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/r02_pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
tail -4 gpurun_out/r02_pytest5.log
V="seg=8;seg=12;seg=16;seg=0"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var5_c3.jsonl 2> gpurun_out/r02_var5_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 64 --check --variants "$V" > gpurun_out/r02_var5_rmat.jsonl 2> gpurun_out/r02_var5_rmat.err
cat gpurun_out/r02_var5_c3.jsonl gpurun_out/r02_var5_rmat.jsonl
timeout 900 python bench.py --workload c3_rmat --coalitions 512 --cpu-coalitions 6 --no-query-leg > gpurun_out/r02_bench_c3_rmat2.json 2> gpurun_out/r02_bench_c3_rmat2.err; echo "rc=$?" >> gpurun_out/r02_bench_c3_rmat2.err
tail -2 gpurun_out/r02_bench_c3_rmat2.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_rmat2.json')); print(d['value'], d['parity']); print(d['cpu_baseline']['y_gpu']); print(d['cpu_baseline']['y_cpu'])"
