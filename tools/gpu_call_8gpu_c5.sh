#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
timeout 500 $TR bench.py --gpus 8 --config c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5_8gpu_fixed.json 2> gpurun_out/r02_bench_c5_8gpu_fixed.err; echo "c5 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_c5_8gpu_fixed.json").read().strip().splitlines()[-1])
    print("c5x8", d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms"], d["elbo"], d["clocks"])
except Exception as e:
    print("c5 failed", e); print(open("gpurun_out/r02_bench_c5_8gpu_fixed.err").read()[-3000:])
PY
