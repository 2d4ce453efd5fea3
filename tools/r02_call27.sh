#!/bin/bash
# round 2, GPU call 27: kernel-level launch list of the C4 step (hetero SAGE, 5 node types / 20 relations)
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_c4.csv python tools/variants.py --workload c4 --coalitions 32 --steps 1 --warmup 1 --variants "seg=8" > gpurun_out/r02_ncu27.log 2>&1
tail -2 gpurun_out/r02_ncu27.log | cut -c1-600
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02_launches_c4.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[vi].replace(",", "")); u = r[ui]
    ms = t / 1e6 if u in ("ns", "nsecond") else (t / 1e3 if u in ("us", "usecond") else t)
    a = agg.setdefault(r[ki][:80], [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
print("total ms", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-82s %4d  %9.3f ms  %5.1f %%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
