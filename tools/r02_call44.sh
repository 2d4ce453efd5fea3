#!/bin/bash
# round 2, GPU call 44: source-level profiles of the layer-0 kernel and of the tile transform (C3)
set -x
CMD="python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants seg=8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"l0_ws_kernel" -s 2 -c 1 -f -o gpurun_out/r02d_l0_ws $CMD > gpurun_out/r02_ncu44a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_tc_kernel" -s 4 -c 1 -f -o gpurun_out/r02d_dense_tc $CMD > gpurun_out/r02_ncu44b.log 2>&1
python tools/ncu_extract.py gpurun_out/r02d_l0_ws.ncu-rep gpurun_out/r02d_dense_tc.ncu-rep > gpurun_out/r02d_l0_dense_ncu.txt
python tools/ncu_lines.py gpurun_out/r02d_l0_ws.ncu-rep 40 > gpurun_out/r02d_l0_ws_lines.txt
python tools/ncu_lines.py gpurun_out/r02d_dense_tc.ncu-rep 40 > gpurun_out/r02d_dense_tc_lines.txt
grep "==\|time_duration\|dram__bytes\|inst_executed\|lsu_wavefronts.avg" gpurun_out/r02d_l0_dense_ncu.txt | cut -c1-140
