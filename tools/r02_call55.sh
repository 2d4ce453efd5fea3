#!/bin/bash
# round 2, GPU call 55: the tile's word of the coalition matrix as a dense column (act_column) at 2048 coalitions (W = 64)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "compact_path or bench_scale or hub or golden or layer0" > gpurun_out/r02_pytest55.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest55.log
timeout 600 python tools/variants.py --workload c3 --coalitions 2048 --steps 1 --warmup 1 --check --variants "act_column=0;act_column=1" > gpurun_out/r02_var55_c3.jsonl 2> gpurun_out/r02_var55_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_var55_c3.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or ({k: round(v, 3) for k, v in d["ms_per_launch"].items()}, round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
