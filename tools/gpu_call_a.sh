#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_a.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_tests_a.log
timeout 300 python tools/gemm_perf.py > gpurun_out/r02_gemm_perf.log 2>&1; echo "gemm rc=$?"; cat gpurun_out/r02_gemm_perf.log | tail -8
timeout 300 python tools/gemm_perf.py 100000 512 >> gpurun_out/r02_gemm_perf.log 2>&1; tail -3 gpurun_out/r02_gemm_perf.log
for cfg in c3 c2 c4 c4f32 c5s; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", d["metric"], d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "kernel_ms", d["roofline"] and d["roofline"]["kernel_ms"])
except Exception as e:
    print("$cfg parse failed", e); print(open("gpurun_out/r02_bench_$cfg.err").read()[-1500:])
PY
done
