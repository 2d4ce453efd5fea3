#!/bin/bash
# round 2, GPU call 49: does the L1 size matter to the segmented SpMM?  shared-memory carve-out 86 % (196 KB, what 6 x 29.9 KB needs) vs 100 % (228 KB)
timeout 900 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --check --variants "seg=8;seg=8,seg_carve=100;seg=8,seg_carve=86;seg=8" > gpurun_out/r02_var49_c3.jsonl 2> gpurun_out/r02_var49_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var49_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
