#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR tools/mgpu_check.py > gpurun_out/r02_mgpu_check.log 2>&1; echo "mgpu rc=$?"; tail -4 gpurun_out/r02_mgpu_check.log
timeout 300 $TR bench.py --gpus 2 --config c5s --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5s_2gpu.json 2> gpurun_out/r02_bench_c5s_2gpu.err; echo "c5s rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c5s_2gpu.json").read().strip().splitlines()[-1])
print("c5s x2", d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms"], d["elbo"])
PY
