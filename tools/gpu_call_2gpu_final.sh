#!/bin/bash
# final 2-GPU check of the round: G = 1 / 2 agreement (tools/mgpu_check.py) and the C3 bench line on 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR tools/mgpu_check.py > gpurun_out/r02_mgpu_check_final.log 2>&1; echo "mgpu rc=$?"; tail -4 gpurun_out/r02_mgpu_check_final.log
timeout 200 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_2gpu_final.json 2> gpurun_out/r02_bench_c3_2gpu_final.err; echo "c3x2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c3_2gpu_final.json").read().strip().splitlines()[-1])
print("c3 x2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["elbo"])
PY
