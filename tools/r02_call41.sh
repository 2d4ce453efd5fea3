#!/bin/bash
# round 2, GPU call 41: R-MAT and C4 with the current segmented kernel (seg_skew 0 / 4), then the full GPU suite
set -x
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg_skew=4;seg_skew=0;seg_skew=6" > gpurun_out/r02_var41_rmat.jsonl 2> gpurun_out/r02_var41_rmat.err
timeout 900 python tools/variants.py --workload c4 --coalitions 64 --steps 1 --warmup 1 --variants "seg=8" > gpurun_out/r02_var41_c4.jsonl 2> gpurun_out/r02_var41_c4.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var41_rmat.jsonl", "gpurun_out/r02_var41_c4.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or ({k: round(v, 3) for k, v in d["ms_per_launch"].items()}, round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest41.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest41.log
