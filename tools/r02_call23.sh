#!/bin/bash
# round 2, GPU call 23: long-row SpMM kernel in (slot, chunk) pass order with a work counter; R-MAT timing + launch list
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "hub or long or rmat or bench_scale or hetero" > gpurun_out/r02_pytest23.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest23.log
tail -3 gpurun_out/r02_pytest23.log
timeout 600 python tools/variants.py --workload c3_rmat --coalitions 128 --check --variants "seg=8;seg=4" > gpurun_out/r02_var23_rmat.jsonl 2> gpurun_out/r02_var23_rmat.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_rmat2.csv python tools/variants.py --workload c3_rmat --coalitions 64 --steps 1 --warmup 1 --variants "seg=8" > gpurun_out/r02_ncu23.log 2>&1
python - <<'PY'
import csv, collections, json
for l in open("gpurun_out/r02_var23_rmat.jsonl"):
    d = json.loads(l)
    print(d.get("variant"), d.get("error") or (round(d["ms_per_launch"]["spmm_tile_l1"], 3), round(d["ms_per_launch"]["spmm_invariant_l0"], 3), round(d["evals_per_s"], 1), d.get("max_rel_diff_vs_first")))
rows = list(csv.reader(l for l in open("gpurun_out/r02_launches_rmat2.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[vi].replace(",", "")); u = r[ui]
    ms = t / 1e6 if u in ("ns", "nsecond") else (t / 1e3 if u in ("us", "usecond") else t)
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:9]:
    print("%-72s %4d  %9.3f ms  %5.1f %%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
