#!/bin/bash
# round 2, GPU call 17 (EXPERIMENTS build): warp-uniform consumer; parity per variant; timing + role counters
set -x
for v in 216 232 316 332 432; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "segmented and $v" > gpurun_out/r02_pytest17_$v.log 2>&1
  echo "variant $v rc=$? $(tail -1 gpurun_out/r02_pytest17_$v.log)"
done
XPGNN_BG_DBG=1 timeout 600 python tools/variants.py --workload c3 --coalitions 128 --check --variants "seg=8;seg=332;seg=432;seg=232;seg=316" > gpurun_out/r02_var17_c3.jsonl 2> gpurun_out/r02_var17_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_var17_c3.jsonl",):
    for l in open(f):
        d = json.loads(l)
        print(f, d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
PY
grep "bg dbg" gpurun_out/r02_var17_c3.err | head -12
