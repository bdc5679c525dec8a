#!/bin/bash
# ncu --set full of the two fast-path pair kernels on the final tree (c3s = C3 block size, 1/16 of the pairs)
mkdir -p gpurun_out
A="--config c3s --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $A > gpurun_out/r02_c3s_plain_final.json 2> gpurun_out/r02_c3s_plain_final.err; rc=$?; echo "plain rc=$rc"
[ $rc -eq 0 ] || exit 1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ks_symx -c 2 -o gpurun_out/r02_symx_full_final python bench.py $A > gpurun_out/r02_ncu_c3s_full_final.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_symx_full_final.ncu-rep
