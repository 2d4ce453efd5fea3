#!/bin/bash
# round 2, GPU call 19: TMA gather4 (UTMALDG.2D.GATHER4) against per-row bulk copies (UBLKCP): rows per second per SM
cd tools
for args in "1 0 4" "4 0 4" "1 0 8" "1 0 16" "4 0 8" "1 1 4" "1 1 8" "1 1 16" "1 0 8 512" "1 0 8 16"; do
  timeout 60 ./tma_gather4_probe $args 2>&1 | tail -1
done | tee ../gpurun_out/r02_tma_gather4_probe.txt
