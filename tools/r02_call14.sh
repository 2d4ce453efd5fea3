#!/bin/bash
# round 2, GPU call 14: warp-specialised SpMM with cp.async producers (seg = 3xx: .cg, 4xx: .ca) against bulk copies (2xx) and seg = 8
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented" > gpurun_out/r02_pytest14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest14.log
tail -6 gpurun_out/r02_pytest14.log
V="seg=8;seg=348;seg=332;seg=316;seg=448;seg=248"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var14_c3.jsonl 2> gpurun_out/r02_var14_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg=8;seg=348" > gpurun_out/r02_var14_rmat.jsonl 2> gpurun_out/r02_var14_rmat.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var14_c3.jsonl", "gpurun_out/r02_var14_rmat.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
tail -3 gpurun_out/r02_var14_c3.err
