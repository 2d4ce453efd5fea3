#!/bin/bash
# round 2, GPU call 40: ncu --set full + source view of the current segmented SpMM (C3)
set -x
CMD="python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants seg=8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cspmm_seg_kernel" -s 2 -c 1 -f -o gpurun_out/r02d_cspmm_seg $CMD > gpurun_out/r02_ncu35.log 2>&1
python tools/ncu_extract.py gpurun_out/r02d_cspmm_seg.ncu-rep > gpurun_out/r02d_cspmm_seg_ncu.txt
python tools/ncu_lines.py gpurun_out/r02d_cspmm_seg.ncu-rep 45 > gpurun_out/r02d_cspmm_seg_lines.txt
grep "time_duration\|xbar2l1tex_read_bytes.sum \|dram__bytes\|hit_rate\|inst_executed\|registers" gpurun_out/r02d_cspmm_seg_ncu.txt | cut -c1-140
head -30 gpurun_out/r02d_cspmm_seg_lines.txt | cut -c1-170
