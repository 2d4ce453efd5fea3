#!/bin/bash
# round 2, GPU call 28: C4 with 4 / 6 / 8 gathers in flight in the segmented SpMM (short rows: ~1.5 active in-edges per relation row)
set -x
timeout 900 python tools/variants.py --workload c4 --coalitions 64 --steps 1 --warmup 1 --check --variants "seg=8;seg=4;seg=6;seg=0" > gpurun_out/r02_var28_c4.jsonl 2> gpurun_out/r02_var28_c4.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var28_c4.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or ({k: round(v, 2) for k, v in d["ms_per_launch"].items()}, d["launches"], round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
tail -2 gpurun_out/r02_var28_c4.err
