#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_final.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; echo "bench c3 rc=$?"
for cfg in c2 c4 c4f32; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err; echo "bench $cfg rc=$?"
done
python - <<'PY'
import json
for cfg in ("c3_1gpu","c2","c4","c4f32"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{cfg}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(cfg, d["metric"], f'{d["value"]:.4g}', "ms", f'{d["ms_per_step"]:.1f}', "e2e", f'{d["e2e"]["value"]:.4g}', "frac", r["frac"], "decl", f'{r["declared_model_ratio"]:.2f}', r["kernel"], r["kernel_instance"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
    except Exception as e:
        print(cfg, "failed", e)
PY
timeout 300 python tools/symx_sweep.py --fwd 0,1 --bwd 0,1 --reps 2 > gpurun_out/r02_sweep_final.log 2>&1; cat gpurun_out/r02_sweep_final.log
timeout 300 python tools/symx_sweep.py --ls 0.3 --fwd 0 --bwd 0 --reps 2 > gpurun_out/r02_sweep_ls03.log 2>&1; cat gpurun_out/r02_sweep_ls03.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__thread_inst_executed_pipe_fp64_pred_on.sum,smsp__thread_inst_executed_pipe_xu_pred_on.sum,smsp__pipe_fp64_cycles_active.sum,smsp__cycles_elapsed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
A="--config c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $A > gpurun_out/r02_plain_c4b.log 2>&1 && ncu --metrics $M --clock-control none -k regex:'ks_fwd_kernel' -c 2 --csv --log-file gpurun_out/r02_ncu_c4.csv python bench.py $A > gpurun_out/r02_ncu_c4b.log 2>&1; echo "ncu c4 rc=$?"
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $B > gpurun_out/r02_plain_c3b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py $B > gpurun_out/r02_ncu_c3b.log 2>&1; echo "ncu launches rc=$? rows=$(wc -l < gpurun_out/r02_launches_c3.csv)"
