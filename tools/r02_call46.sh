#!/bin/bash
# round 2, GPU call 46: position-parallel id staging (skewed blocks) with four list loads in flight per lane; R-MAT and C3
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale or hub" > gpurun_out/r02_pytest46.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest46.log
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --variants "seg=8;seg=8" > gpurun_out/r02_var46_rmat.jsonl 2> gpurun_out/r02_var46_rmat.err
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --variants "seg=8" > gpurun_out/r02_var46_c3.jsonl 2> gpurun_out/r02_var46_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var46_rmat.jsonl", "gpurun_out/r02_var46_c3.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or ({k: round(v, 3) for k, v in d["ms_per_launch"].items()}, round(d["evals_per_s"], 1)))
PY
