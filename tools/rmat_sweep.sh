#!/bin/bash
# short R-MAT C3 bench: per-kernel ms per launch
out=$(env "$@" timeout 300 python bench.py --workload c3_rmat --steps 1 --warmup 3 --no-cpu-baseline --coalitions-per-gpu 64 2>gpurun_out/rmat.err)
python - "$out" <<PY
import json, sys
try:
    d = json.loads(sys.argv[1])
    print("rmat | evals/s %.0f |" % d["value"], {k: round(v["ms"] / max(v["launches"], 1), 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e, sys.argv[1][-300:]); print(open("gpurun_out/rmat.err").read()[-800:])
PY
