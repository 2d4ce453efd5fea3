#!/bin/bash
# round 2, GPU call 16 (EXPERIMENTS build): per-role cycle counters of the warp-specialised SpMM
set -x
XPGNN_BG_DBG=1 timeout 600 python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants "seg=348;seg=248" > gpurun_out/r02_var16_c3.jsonl 2> gpurun_out/r02_var16_c3.err
grep "bg dbg" gpurun_out/r02_var16_c3.err | head -12
