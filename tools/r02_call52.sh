#!/bin/bash
# round 2, GPU call 52: default seg = 7 (7 gathers in flight, 7 CTAs per SM, 25.9 KB of shared memory per CTA): tests, C3, R-MAT, C4
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale or hub or hetero" > gpurun_out/r02_pytest52.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest52.log
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg=7;seg=8;seg=8,seg_occ=7" > gpurun_out/r02_var52_rmat.jsonl 2> gpurun_out/r02_var52_rmat.err
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --variants "seg=7" > gpurun_out/r02_var52_c3.jsonl 2> gpurun_out/r02_var52_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var52_rmat.jsonl", "gpurun_out/r02_var52_c3.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
