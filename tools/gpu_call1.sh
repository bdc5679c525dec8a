#!/bin/bash
# r02 call 1: fast-path sweep at C3, extended issue model, GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c1_smi.txt 2>&1
timeout 600 python tools/symx_sweep.py > gpurun_out/r02_sweep1.log 2>&1; echo "sweep rc=$?"
timeout 120 ./tools/issue_model > gpurun_out/r02_issue_model.txt 2>&1; echo "issue rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests1.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_tests1.log
cat gpurun_out/r02_sweep1.log
