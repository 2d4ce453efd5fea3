#!/bin/bash
# round 2, GPU call 22: full GPU test suite on the current tree; kernel-level launch list of the R-MAT C3 step
set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest22.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest22.log
tail -5 gpurun_out/r02_pytest22.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_rmat.csv python tools/variants.py --workload c3_rmat --coalitions 64 --steps 1 --warmup 1 --variants "seg=8" > gpurun_out/r02_ncu22.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02_launches_rmat.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[vi].replace(",", "")); u = r[ui]
    ms = t / 1e6 if u in ("ns", "nsecond") else (t / 1e3 if u in ("us", "usecond") else t)
    k = r[ki][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%-72s %4d  %9.3f ms  %5.1f %%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
