#!/bin/bash
# round 2, GPU call 26: the driver-style bench line (default arguments) and its reference arm; the R-MAT line
set -x
timeout 900 python bench.py > gpurun_out/r02_bench_c3_final.json 2> gpurun_out/r02_bench_c3_final.err
echo "rc=$?" >> gpurun_out/r02_bench_c3_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm_final.json 2> gpurun_out/r02_bench_reference_arm_final.err
echo "rc=$?" >> gpurun_out/r02_bench_reference_arm_final.err
timeout 900 python bench.py --workload c3_rmat --coalitions 1024 --no-query-leg > gpurun_out/r02_bench_c3_rmat_final.json 2> gpurun_out/r02_bench_c3_rmat_final.err
echo "rc=$?" >> gpurun_out/r02_bench_c3_rmat_final.err
tail -2 gpurun_out/r02_bench_c3_final.err gpurun_out/r02_bench_reference_arm_final.err gpurun_out/r02_bench_c3_rmat_final.err | cut -c1-300
python - <<'PY'
import json
for f in ("r02_bench_c3_final", "r02_bench_reference_arm_final", "r02_bench_c3_rmat_final"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("e2e", {}).get("value"), d.get("parity"), d.get("clocks"))
        q = d.get("s_per_explained_query")
        if q: print({k: (v.get("warm_s") if isinstance(v, dict) else v) for k, v in q.items()})
    except Exception as ex:
        print(f, "ERR", ex)
PY
