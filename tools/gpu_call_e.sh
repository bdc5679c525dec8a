#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_e.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_e.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for ls in 0.3 0.25 0.2; do timeout 200 python tools/symx_sweep.py --ls $ls --fwd 0 --bwd 0 --reps 2 2>&1 | sed "s/^/ls=$ls /"; done | tee gpurun_out/r02_sweep_clamp.log
