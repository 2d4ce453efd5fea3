#!/bin/bash
# round 2, GPU call 21: hybrid launch with the gather4 kernel on a high-priority stream and one segmented CTA less per SM
set -x
V="seg=8;seg=8,seg_tma=1;seg=8,seg_occ=7,seg_tma=1;seg=6,seg_occ=6,seg_tma=1"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var21_c3.jsonl 2> gpurun_out/r02_var21_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var21_c3.jsonl",):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
tail -3 gpurun_out/r02_var21_c3.err
