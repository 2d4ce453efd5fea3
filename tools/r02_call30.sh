#!/bin/bash
# round 2, GPU call 30 (2 GPUs): smoke(), and the default bench under torchrun on 2 GPUs (driver-style)
set -x
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke30.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke30.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_2gpu_final.json 2> gpurun_out/r02_bench_c3_2gpu_final.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_2gpu_final.json 2> gpurun_out/r02_bench_ref_2gpu_final.err
echo "ref rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_c3_2gpu_final", "r02_bench_ref_2gpu_final"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("n_gpus"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("e2e", {}).get("value"), d.get("ms_per_step_by_rank"), d.get("clocks"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
