#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_c.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_tests_c.log
timeout 300 python tools/gemm_perf.py > gpurun_out/r02_gemm_perf.log 2>&1; echo "gemm rc=$?"; tail -4 gpurun_out/r02_gemm_perf.log
timeout 300 python tools/gemm_perf.py 100000 512 >> gpurun_out/r02_gemm_perf.log 2>&1; tail -3 gpurun_out/r02_gemm_perf.log
