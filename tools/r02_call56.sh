#!/bin/bash
# round 2, GPU call 56: full GPU suite and smoke on the final tree
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest56.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest56.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke56.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke56.log
