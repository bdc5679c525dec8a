#!/bin/bash
mkdir -p gpurun_out
./tools/rf_probe > gpurun_out/r02_rf_probe.txt 2>&1; echo "rf rc=$?"
cat gpurun_out/r02_rf_probe.txt
timeout 600 python tools/symx_sweep.py --fwd 5 > gpurun_out/r02_sweep2.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/r02_sweep2.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests2.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02_tests2.log
