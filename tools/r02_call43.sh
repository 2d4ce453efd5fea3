#!/bin/bash
# round 2, GPU call 43: list loads of the staging loop with the evict-first hint (ld.global.cs)
timeout 900 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --variants "seg=8;seg=8" > gpurun_out/r02_var43_c3.jsonl 2> gpurun_out/r02_var43_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var43_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1)))
PY
