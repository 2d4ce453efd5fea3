#!/bin/bash
# round 2, GPU call 50: segmented SpMM with 25.9 KB of shared memory per CTA (448 staged positions, unpadded tile rows): 164 KB carve-out, 92 KB L1
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale or hub" > gpurun_out/r02_pytest50.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest50.log
timeout 900 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --check --variants "seg=8;seg=8" > gpurun_out/r02_var50_c3.jsonl 2> gpurun_out/r02_var50_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --variants "seg=8" > gpurun_out/r02_var50_rmat.jsonl 2> gpurun_out/r02_var50_rmat.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var50_c3.jsonl", "gpurun_out/r02_var50_rmat.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
