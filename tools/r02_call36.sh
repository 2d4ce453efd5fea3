#!/bin/bash
# round 2, GPU call 36: segmented SpMM, staging buffer of 512 positions (one round per C3 block)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale" > gpurun_out/r02_pytest36.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest36.log
timeout 600 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --variants "seg=8" > gpurun_out/r02_var36_c3.jsonl 2> gpurun_out/r02_var36_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --variants "seg=8" > gpurun_out/r02_var36_rmat.jsonl 2> gpurun_out/r02_var36_rmat.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var36_c3.jsonl", "gpurun_out/r02_var36_rmat.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or ({k: round(v, 3) for k, v in d["ms_per_launch"].items()}, round(d["evals_per_s"], 1)))
PY
