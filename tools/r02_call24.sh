#!/bin/bash
# round 2, GPU call 24: sliced hub rows in layer 0; R-MAT timing
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "hub or layer0 or bench_scale or fp32_exact or rmat" > gpurun_out/r02_pytest24.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest24.log
tail -3 gpurun_out/r02_pytest24.log
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "l0_slices=0;l0_slices=1" > gpurun_out/r02_var24_rmat.jsonl 2> gpurun_out/r02_var24_rmat.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var24_rmat.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["ms_per_launch"]["spmm_invariant_l0"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
tail -2 gpurun_out/r02_var24_rmat.err
