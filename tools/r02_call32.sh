#!/bin/bash
# round 2, GPU calls 32 / 33: segmented SpMM with the row scale computed once per row (call 32: 11.12 -> 10.77 ms per C3 tile), then
# with the accumulator reset folded into the FMA on top of it (call 33: 10.79 ms, no gain, reverted)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale" > gpurun_out/r02_pytest32.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest32.log
timeout 600 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --variants "seg=8" > gpurun_out/r02_var32_c3.jsonl 2> gpurun_out/r02_var32_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --variants "seg=8" > gpurun_out/r02_var32_rmat.jsonl 2> gpurun_out/r02_var32_rmat.err
