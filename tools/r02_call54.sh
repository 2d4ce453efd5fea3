#!/bin/bash
# round 2, GPU call 54: C4 (hetero SAGE) bench line with the final build
timeout 800 python bench.py --workload c4 --coalitions 128 --cpu-coalitions 1 --no-query-leg > gpurun_out/r02_bench_c4_final.json 2> gpurun_out/r02_bench_c4_final.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c4_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("parity"), d["clocks"])
print({k: (round(v["ms"] / v["launches"], 3), v["launches"]) for k, v in d["kernels"].items() if v["launches"]})
PY
