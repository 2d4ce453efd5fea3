#!/bin/bash
# round 2, GPU call 20: TMA gather4 producers (seg = 516 / 532 standalone) and the hybrid launch (seg_tma = 1: gather4 kernel next to the segmented kernel)
set -x
for v in 516 532 8t 6t; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "segmented and $v" > gpurun_out/r02_pytest20_$v.log 2>&1
  echo "variant $v rc=$? $(tail -1 gpurun_out/r02_pytest20_$v.log)"
done
V="seg=8;seg=8,seg_tma=1;seg=6,seg_tma=1;seg=4,seg_tma=1;seg=532;seg=516"
timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "$V" > gpurun_out/r02_var20_c3.jsonl 2> gpurun_out/r02_var20_c3.err
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg=8;seg=8,seg_tma=1;seg=4,seg_tma=1" > gpurun_out/r02_var20_rmat.jsonl 2> gpurun_out/r02_var20_rmat.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var20_c3.jsonl", "gpurun_out/r02_var20_rmat.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
tail -3 gpurun_out/r02_var20_c3.err
