#!/bin/bash
# round 2, GPU call 47: segmented SpMM, piece boundaries at the row start nearest to g * len / 4
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale or hub" > gpurun_out/r02_pytest47.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest47.log
timeout 600 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --variants "seg=8;seg=8" > gpurun_out/r02_var47_c3.jsonl 2> gpurun_out/r02_var47_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var47_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1)))
PY
