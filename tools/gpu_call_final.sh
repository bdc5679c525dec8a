#!/bin/bash
# final verification of the round: GPU test suite, smoke, default bench line, reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_final_ref.json 2>&1; echo "ref rc=$?"; tail -c 800 gpurun_out/r02_bench_final_ref.json
