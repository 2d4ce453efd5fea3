#!/bin/bash
# round 2, GPU call 57: the driver-style bench line of the final tree (default arguments)
timeout 400 python bench.py > gpurun_out/r02_bench_c3_final.json 2> gpurun_out/r02_bench_c3_final.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c3_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["ok"], d["clocks"])
print({k: round(v["ms"] / v["launches"], 3) for k, v in d["kernels"].items() if v["launches"]})
q = d.get("s_per_explained_query") or {}
print({k: (round(v.get("warm_s"), 3) if isinstance(v, dict) and "warm_s" in v else None) for k, v in q.items()})
PY
