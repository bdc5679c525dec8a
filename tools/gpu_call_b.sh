#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_b.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_tests_b.log
timeout 300 python tools/gemm_perf.py > gpurun_out/r02_gemm_perf.log 2>&1; echo "gemm rc=$?"; tail -4 gpurun_out/r02_gemm_perf.log
timeout 300 python tools/gemm_perf.py 200000 1024 > gpurun_out/r02_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gram_syrk|gemm_wm' -c 3 -o gpurun_out/r02_gemm python tools/gemm_perf.py 200000 1024 > gpurun_out/r02_gemm_ncu.log 2>&1; echo "ncu rc=$?"
