#!/bin/bash
# round 2, GPU call 37: gathers in flight / occupancy of the segmented SpMM re-swept with the 512-position staging buffer
timeout 900 python tools/variants.py --workload c3 --coalitions 128 --check --variants "seg=8;seg=8,seg_occ=7;seg=6;seg=6,seg_occ=7;seg=12;seg=4;seg=0" > gpurun_out/r02_var37_c3.jsonl 2> gpurun_out/r02_var37_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var37_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
