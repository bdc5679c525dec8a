#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_8gpu_smi.txt
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_8gpu.json 2> gpurun_out/r02_bench_c3_8gpu.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c3_8gpu.json").read().strip().splitlines()[-1])
print("c3x8", d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["kernel_ms"], d["elbo"])
PY
timeout 900 $TR bench.py --gpus 8 --config c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c5_8gpu.json 2> gpurun_out/r02_bench_c5_8gpu.err; echo "c5 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_c5_8gpu.json").read().strip().splitlines()[-1])
    print("c5x8", d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["kernel_ms"], d["elbo"], d["config"])
except Exception as e:
    print("c5 failed", e); print(open("gpurun_out/r02_bench_c5_8gpu.err").read()[-3000:])
PY
