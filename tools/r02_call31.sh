#!/bin/bash
# round 2, GPU call 31: segmented SpMM at 10 CTAs per SM with 4 gathers in flight (40 warps, <= 51 registers)
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "seg=8;seg=4,seg_occ=10;seg=4;seg=6,seg_occ=8" > gpurun_out/r02_var31_c3.jsonl 2> gpurun_out/r02_var31_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var31_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
