#!/usr/bin/env python
"""Benchmark of the CaGP hot path: ELBO+gradient steps per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # reference-equivalent CPU path

One "step" = value_and_grad of -ELBO(init(train_data)) w.r.t. {lengthscale, variance, obs_stddev,
mean constant, nz_values} on synthetic data of the named shape (config C3 by default:
n = 1M, d = 3, RBF, BlockSparsePolicy k = 1024, float64).  N > 1 shards the rows of X over the
ranks (one process per GPU, torchrun); total work is fixed, so scaling is "strong".

Prints ONE JSON line (rank 0).  Keys: the contract of the build spec plus `roofline` (fp64 pipe),
`cpu_baseline`, `e2e`, `gpu_launches`, `clocks`, `kernel_evals_per_sec`.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n, d, family, k, dtype)
    "c1": dict(n=1000, d=1, family="rbf", k=10, dtype="f64", desc="README example shape n=1000,d=1,RBF,k=10"),
    "c2": dict(n=100_000, d=8, family="matern32", k=512, dtype="f64", desc="n=100k,d=8,Matern-3/2,k=512"),
    # profiling-size variant of c3: same block size (976 rows per action block), 1/16 of the pairs
    "c3s": dict(n=250_000, d=3, family="rbf", k=256, dtype="f64", desc="n=250k,d=3,RBF,BlockSparse k=256,float64 (c3 block size, ncu-sized)"),
    # BASELINE config 5 shape (dense Gaussian action matrix, d = 16, k = 256) at a size one step of which
    # takes seconds: n = 200k (the named n = 2M is 100x the pairs)
    "c5s": dict(n=200_000, d=16, family="rbf", k=256, dtype="f64", dense=True, lengthscale=4.0,
                desc="n=200k,d=16,RBF,dense Gaussian actions k=256,float64 (config-5 shape, reduced n)"),
    "c3": dict(n=1_000_000, d=3, family="rbf", k=1024, dtype="f64", desc="n=1M,d=3,RBF,BlockSparse k=1024,float64"),
}

# Declared algorithmic flops per kernel pair (SURVEY.md 8d / BASELINE.md 3): FMA = 2, fp64 exp = 31.
def pair_flops(family: str, d: int, backward: bool) -> int:
    base = {"rbf": 3 * d + 34, "matern12": 3 * d + 51, "matern32": 3 * d + 53, "matern52": 3 * d + 55}[family]
    return base + (6 if backward else 0)


# Measured fp64 issue rate on this pool's B200 (tools/fp64_ops.cu, profiles/r01_issue_model.txt): DADD, DMUL and
# DADD/DMUL + DFMA mixes - what the pair kernels execute - issue at 18.56 T lane-instructions/s = 37.1 TFLOP/s
# FMA-equivalent (nominal 148 x 64 x 2 x 1.965 GHz = 37.2); a pure DFMA stream settles at 34.2.  MEASURED_PEAKS.json
# has no fp64 entry, so this measured number is the denominator (the stricter of the two).
FP64_PEAK_TFLOPS = 37.1

# fp64 issue slots actually executed per ORDERED pair by each pair kernel (counted in the SASS of the
# d = 3 RBF instances; the symmetric kernels evaluate each unordered pair once = half per ordered pair, and
# run the expanded-distance form on this workload: 12 slots per unordered pair forward, 16 backward).
FP64_SLOTS_PER_PAIR = {"cagp_ks_blocksparse_fwd": 14.0, "cagp_ks_blocksparse_bwd": 16.0,
                       "cagp_ks_blocksparse_fwd_sym": 6.0, "cagp_ks_blocksparse_fwd_sym_peer": 6.0,
                       "cagp_ks_blocksparse_bwd_sym": 8.0}


def synth(cfg, device, dtype):
    """Deterministic synthetic inputs of the named shape (SURVEY 8d seeds for C3)."""
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    g = torch.Generator().manual_seed(11)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 10.0
    if cfg.get("dense"):
        x = torch.randn(n, d, generator=torch.Generator().manual_seed(21), dtype=torch.float64)
    g = torch.Generator().manual_seed(12)
    y = torch.sin(x[:, 0]) * torch.cos(x[:, min(1, d - 1)]) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    g = torch.Generator().manual_seed(13)
    if cfg.get("dense"):
        s = torch.randn(n, k, generator=torch.Generator().manual_seed(22), dtype=torch.float64) / n**0.5
    else:
        s = torch.randn(n, generator=g, dtype=torch.float64)
    return x.to(dtype), y.to(dtype), s.to(dtype)


def build_model(cfg, x, y, s, device, dtype, process_group=None):
    import cagpjax_b200 as cagp
    from cagpjax_b200 import gpx
    from cagpjax_b200.policies.block_sparse import _normalize_by_blocks

    kern_cls = {"rbf": gpx.RBF, "matern12": gpx.Matern12, "matern32": gpx.Matern32, "matern52": gpx.Matern52}[cfg["family"]]
    t = lambda v: torch.tensor(v, dtype=dtype, device=device)
    kernel = kern_cls(lengthscale=t(cfg.get("lengthscale", 1.0)), variance=t(1.0),
                      compute_engine=cagp.computations.LazyKernelComputation(process_group=process_group))
    prior = gpx.Prior(kernel=kernel, mean_function=gpx.Constant(t(0.0)))
    lik = gpx.Gaussian(cfg["n"], obs_stddev=t(0.1))
    if cfg.get("dense"):
        policy = cagp.policies.DensePolicy(s.to(device))
    else:
        policy = cagp.BlockSparsePolicy(cfg["k"], _normalize_by_blocks(s, cfg["k"]).to(device))
    model = cagp.ComputationAwareGP(prior * lik, policy)
    data = gpx.Dataset(x.to(device), y.to(device).reshape(-1, 1))
    return model, data


def neg_elbo(model, data):
    return -model.elbo(model.init(data))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        return False

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_step_seconds(cfg, n_sample: int, threads: int):
    """Times ONE ELBO+grad step of the oracle's literal restatement of the reference path
    (row-batched dense K chunk @ dense S, lazy_kernel.py:79-90) at n = n_sample on the host."""
    from oracle import cagp_oracle as O

    torch.set_num_threads(threads)
    cfg_s = dict(cfg, n=n_sample)
    k = max(1, min(cfg["k"], n_sample // 8))
    x, y, s = synth(cfg_s, "cpu", torch.float64)
    s = O.normalize_by_blocks(s, k)
    p = O.make_params(1.0, 1.0, 0.1, 0.0, s)
    t0 = time.perf_counter()
    O.elbo_and_grad_literal(cfg["family"], x, y, p, k)
    return time.perf_counter() - t0, k


def cpu_structured_pairs_per_sec(cfg, n_sample: int = 12288):
    """Kernel pairs per second of the oracle's STRUCTURED O(n^2) forward K'S (never densifies S) on the host
    (lets the algorithmic speed-up - no dense n x k products - be separated from the hardware one)."""
    from oracle import cagp_oracle as O

    cfg_s = dict(cfg, n=n_sample)
    k = max(1, min(cfg["k"], n_sample // 8))
    x, y, s = synth(cfg_s, "cpu", torch.float64)
    t0 = time.perf_counter()
    O.projected_kernel_blocksparse(cfg["family"], x.numpy(), x.numpy(), np.ones(1), s.numpy(), k)
    return n_sample * n_sample / (time.perf_counter() - t0)


def run_reference(args, cfg):
    """--impl reference: the reference's own implementation cannot be imported (no jax / gpjax / cola in
    this image, no network), so this times the oracle port of its algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.ref_n
    times = []
    for i in range(args.warmup + args.steps):
        dt, k_s = cpu_port_step_seconds(cfg, n_sample, threads)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    # the reference algorithm is O(n^2 k): scale the bounded sample to the named shape
    scale = (cfg["n"] / n_sample) ** 2 * (cfg["k"] / k_s)
    t_full = t * scale
    value = 1.0 / t_full
    sample = f"oracle literal port at n={n_sample},k={k_s} ({t:.2f} s/step measured), extrapolated x{scale:.3g} by n^2*k to n={cfg['n']},k={cfg['k']}"
    out = {
        "impl": "reference", "metric": "elbo_grad_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "n": cfg["n"], "d": cfg["d"], "k": cfg["k"], "kernel": cfg["family"]},
        "kernel_evals_per_sec": 2.0 * cfg["n"] ** 2 * value,
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--ref-n", type=int, default=6144, help="sample size of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]

    if args.impl == "reference":
        run_reference(args, cfg)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for this path")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    if world != args.gpus:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    from cagpjax_b200 import _native, gpx

    dtype = torch.float64 if cfg["dtype"] == "f64" else torch.float32
    x, y, s = synth(cfg, device, dtype)
    model, data = build_model(cfg, x, y, s, device, dtype, group)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return gpx.value_and_grad(neg_elbo, model, data)

    for _ in range(args.warmup):
        step()
    barrier()
    _native.PROFILE = []
    launches0 = _native.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            val, grads = step()
        e1.record()
        barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = _native.LAUNCHES - launches0
    profile, _native.PROFILE = _native.PROFILE, None
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = 1e3 / ms_per_step

    # per-kernel device times (CUDA events on the launching stream, inside the timed region)
    per_kernel = {}
    for name, a, b, work in profile:
        d = per_kernel.setdefault(name, [0.0, 0, 0])
        d[0] += a.elapsed_time(b); d[1] += 1; d[2] += work
    n, dd, k = cfg["n"], cfg["d"], cfg["k"]
    fam = cfg["family"]
    kern_ms = {nm: v[0] / v[1] for nm, v in per_kernel.items()}
    dom = max((nm for nm in kern_ms if "ks_" in nm), key=lambda nm: kern_ms[nm])
    pairs_per_launch = per_kernel[dom][2] / per_kernel[dom][1]
    flops = pairs_per_launch * (pair_flops(fam, dd, "bwd" in dom) + (2.0 * k if "dense" in dom else 0.0))
    achieved = flops / (kern_ms[dom] * 1e-3) / 1e12
    roofline = {"bound": "fp64", "kernel": dom, "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / FP64_PEAK_TFLOPS,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from an ncu
                # capture of this command at N = 1 (profiles/r01_traffic_c3_sym.csv); algorithmic: G' = 8 n k bytes
                "traffic": 8.84e9 if (args.config == "c3" and world == 1 and dom == "cagp_ks_blocksparse_bwd_sym") else None,
                "note": "achieved = declared algorithmic flops (libm-exp cost model, SURVEY 8d: 3d+34 fwd / 3d+40 bwd per ordered pair) / CUDA-event kernel time; peak = measured fp64 issue rate of this pool's B200 (37.1 TFLOP/s FMA-equivalent for the DADD/DMUL/DFMA mix; pure DFMA 34.2; nominal 37.2). frac > 1 because the kernels spend fewer fp64 slots than the declared model (table-based exp2, symmetric tiles); fp64_pipe_frac = executed fp64 slots x 2 / time / peak is the pipe utilisation",
                "pairs_per_sec": pairs_per_launch / (kern_ms[dom] * 1e-3),
                "fp64_pipe_frac": (pairs_per_launch / (kern_ms[dom] * 1e-3)) * FP64_SLOTS_PER_PAIR.get(dom, 0.0) * 2.0 / 1e12 / FP64_PEAK_TFLOPS
                if (fam == "rbf" and dd == 3 and cfg["dtype"] == "f64") else None,
                "kernel_ms": {nm: round(v, 3) for nm, v in kern_ms.items()}}

    # end-to-end: host buffers -> device, step, gradients -> host, every step
    e2e = None
    if not args.no_e2e:
        hx, hy, hs = (t.cpu().pin_memory() for t in (x, y, s))
        h2d = sum(t.numel() * t.element_size() for t in (hx, hy, hs))
        d2h = 0
        times = []
        for i in range(1 + max(1, args.steps // 2)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dx, dy, ds = (t.to(device, non_blocking=True) for t in (hx, hy, hs))
            m2, d2 = build_model(cfg, dx, dy, ds, device, dtype, group)
            val, grads = gpx.value_and_grad(neg_elbo, m2, d2)
            host = [val.cpu()] + [g.cpu() for g in grads.values()]
            torch.cuda.synchronize()
            if i > 0:
                times.append(time.perf_counter() - t0)
            d2h = sum(t.numel() * t.element_size() for t in host)
        te = sum(times) / len(times)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([te], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": 1.0 / te, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dt, k_s = cpu_port_step_seconds(cfg, args.ref_n, threads)
        dt, k_s = cpu_port_step_seconds(cfg, args.ref_n, threads)
        scale = (n / args.ref_n) ** 2 * (k / k_s)
        cpu_baseline = {"value": 1.0 / (dt * scale), "unit": "steps/s", "cores": threads, "kind": "port",
                        "sample": f"oracle literal port of the reference path at n={args.ref_n},k={k_s}: {dt:.2f} s/step measured, extrapolated x{scale:.3g} (n^2 k) to the named shape",
                        "structured_fwd_pairs_per_sec": cpu_structured_pairs_per_sec(cfg)}

    if rank == 0:
        out = {
            "metric": "elbo_grad_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": cfg["desc"], "n": n, "d": dd, "k": k, "kernel": fam, "parallelism": f"rows sharded x{world}",
                       "l2": "per-step working set (M' = 8*n*k bytes) exceeds L2; no explicit flush"},
            "kernel_evals_per_sec": 2.0 * n * n * value,
            "elbo": float(-val), "gpu_launches": launches, "clocks": clocks.summary(),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        }
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
