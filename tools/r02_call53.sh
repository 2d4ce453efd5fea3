#!/bin/bash
# round 2, GPU call 53: full GPU suite + smoke on the final build; ncu --set full + source lines of the shipped SpMM (seg = 7)
set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest53.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest53.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke53.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke53.log
CMD="python tools/variants.py --workload c3 --coalitions 64 --steps 1 --warmup 1 --variants seg=7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cspmm_seg_kernel" -s 2 -c 1 -f -o gpurun_out/r02e_cspmm_seg $CMD > gpurun_out/r02_ncu53.log 2>&1
python tools/ncu_extract.py gpurun_out/r02e_cspmm_seg.ncu-rep > gpurun_out/r02e_cspmm_seg_ncu.txt
python tools/ncu_lines.py gpurun_out/r02e_cspmm_seg.ncu-rep 30 > gpurun_out/r02e_cspmm_seg_lines.txt
grep "==\|time_duration\|xbar2l1tex_read_bytes\|dram__bytes\|hit_rate\|inst_executed\|registers\|warps_active\|occupancy_limit" gpurun_out/r02e_cspmm_seg_ncu.txt | cut -c1-140
