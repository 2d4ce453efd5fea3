#!/bin/bash
# round 2, GPU call 39: segmented SpMM, list entries of the next block prefetched during the epilogue (seg_pf = 0 / 8 / 12 / 16)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "segmented or compact_path or bench_scale" > gpurun_out/r02_pytest39.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest39.log
timeout 900 python tools/variants.py --workload c3 --coalitions 256 --steps 2 --warmup 2 --check --variants "seg_pf=0;seg_pf=8;seg_pf=12;seg_pf=16;seg_pf=0;seg_pf=8" > gpurun_out/r02_var39_c3.jsonl 2> gpurun_out/r02_var39_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg_pf=0;seg_pf=8" > gpurun_out/r02_var39_rmat.jsonl 2> gpurun_out/r02_var39_rmat.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var39_c3.jsonl", "gpurun_out/r02_var39_rmat.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
