#!/bin/bash
# round 2, GPU call 18 (EXPERIMENTS build): 16 consumer warps; ceiling experiment with pseudo-random ids (no search / list load in the producers)
set -x
XPGNN_BG_DBG=1 timeout 600 python tools/variants.py --workload c3 --coalitions 128 --variants "seg=8;seg=332;seg=332,l2_gather=77;seg=232,l2_gather=77" > gpurun_out/r02_var18_c3.jsonl 2> gpurun_out/r02_var18_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var18_c3.jsonl",):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1)))
PY
grep "bg dbg" gpurun_out/r02_var18_c3.err | awk 'NR%2==1' | head -12
