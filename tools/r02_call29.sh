#!/bin/bash
# round 2, GPU call 29: ncu --set full of the current dense_tc kernel (C3) and of the long-row kernels (R-MAT C3), text exports
set -x
C3="python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants seg=8"
RM="python tools/variants.py --workload c3_rmat --coalitions 64 --steps 1 --warmup 1 --variants seg=8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_tc_kernel" -s 3 -c 1 -f -o gpurun_out/r02b_dense_tc $C3 > gpurun_out/r02_ncu29a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cspmm_long_kernel" -s 1 -c 1 -f -o gpurun_out/r02b_cspmm_long $RM > gpurun_out/r02_ncu29b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"l0_rows_kernel" -s 1 -c 1 -f -o gpurun_out/r02b_l0_long $RM > gpurun_out/r02_ncu29c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cspmm_seg_kernel" -s 1 -c 1 -f -o gpurun_out/r02b_cspmm_seg_rmat $RM > gpurun_out/r02_ncu29d.log 2>&1
python tools/ncu_extract.py gpurun_out/r02b_dense_tc.ncu-rep gpurun_out/r02b_cspmm_long.ncu-rep gpurun_out/r02b_l0_long.ncu-rep gpurun_out/r02b_cspmm_seg_rmat.ncu-rep > gpurun_out/r02b_ncu.txt
grep -c "==" gpurun_out/r02b_ncu.txt; grep "==\|time_duration\|dram__bytes\|xbar2l1tex_read_bytes.sum \|hit_rate\|lsu_wavefronts.avg\|long_scoreboard\|issue_active" gpurun_out/r02b_ncu.txt | cut -c1-150
