#!/bin/bash
# round 2, GPU call 58: hub-row tests after capping the slice scratch
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "hub or layer0 or bench_scale or golden" > gpurun_out/r02_pytest58.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest58.log
