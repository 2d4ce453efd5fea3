#!/bin/bash
# usage: tools/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- runs a short C3 bench per environment and prints per-kernel ms per launch
for cfg in "$@"; do
  out=$(env $cfg timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --coalitions-per-gpu ${SWEEP_COALITIONS:-64} ${SWEEP_ARGS:-} 2>gpurun_out/sweep.err)
  python - "$cfg" "$out" <<PY
import json, sys
cfg, out = sys.argv[1], sys.argv[2]
try:
    d = json.loads(out)
    print(cfg, "| evals/s %.0f |" % d["value"], {k: round(v["ms"] / max(v["launches"], 1), 2) for k, v in d["kernels"].items()}, "frac %.3f" % (d["roofline"]["frac"] or 0))
except Exception as e:
    print(cfg, "FAILED", e, out[-300:])
PY
done
