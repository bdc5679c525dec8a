#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_d.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_tests_d.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_c3_d.json 2> gpurun_out/r02_bench_c3_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c3_d.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["elbo"])
PY
