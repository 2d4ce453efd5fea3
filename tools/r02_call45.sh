#!/bin/bash
# round 2, GPU call 45: waiting warps of dense_tc / l0_ws suspend inside mbarrier.try_wait (suspend-time hint) instead of polling
V="dense_wait_ns=0,l0_wait_ns=0;dense_wait_ns=500;dense_wait_ns=2000;dense_wait_ns=20000;l0_wait_ns=500;l0_wait_ns=2000;l0_wait_ns=20000;dense_wait_ns=2000,l0_wait_ns=2000"
timeout 900 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var45_c3.jsonl 2> gpurun_out/r02_var45_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var45_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or ({k: round(v, 3) for k, v in d["ms_per_launch"].items() if k in ("spmm_invariant_l0", "dense", "spmm_tile_l1")}, round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
tail -2 gpurun_out/r02_var45_c3.err
