#!/bin/bash
# round 2, GPU call 51: 7 CTAs per SM with 25.9 KB of shared memory per CTA (7 fit the 196 KB carve-out: 60 KB of L1 stay)
timeout 900 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --check --variants "seg=8;seg=7;seg=8,seg_occ=7;seg=6,seg_occ=7;seg=6,seg_occ=6;seg=8" > gpurun_out/r02_var51_c3.jsonl 2> gpurun_out/r02_var51_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var51_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
