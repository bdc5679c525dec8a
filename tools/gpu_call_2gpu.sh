#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_sharding_gpu.py -x -q 2>&1 | tail -3
timeout 300 $TR tools/mgpu_check.py > gpurun_out/r02_mgpu_check.log 2>&1; echo "mgpu rc=$?"; tail -9 gpurun_out/r02_mgpu_check.log
timeout 300 $TR tools/sanitize_peer.py > gpurun_out/r02_peer_plain.log 2>&1; echo "peer rc=$?"; tail -5 gpurun_out/r02_peer_plain.log
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_c3_2gpu.json 2> gpurun_out/r02_bench_c3_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c3_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["kernel_ms"], d["elbo"])
PY
tail -3 gpurun_out/r02_bench_c3_2gpu.err
