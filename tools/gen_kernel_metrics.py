"""profiles/kernel_metrics.json from the ncu csv files of tools/ncu_kernel_metrics.sh (committed as
profiles/r02_ncu_<cfg>.csv): per bench config and C entry point, the executed floating-point instructions per
ALGORITHMIC ordered pair, the fp64-pipe active share, DRAM bytes and duration per launch of the kernel instance that
does the work.  bench.py multiplies fp_inst_per_pair by its live pair rate to report the executed-instruction roofline.
usage: gen_kernel_metrics.py profiles/r02_ncu_c3.csv [more csv ...]   (config name = file name suffix)"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench as B

ENTRY = [(r"ks_symx?_fwd_kernel", "cagp_ks_blocksparse_fwd_sym"), (r"ks_symx?_bwd_kernel", "cagp_ks_blocksparse_bwd_sym"),
         (r"ks_fwd_kernel", "cagp_ks_blocksparse_fwd"), (r"ks_bwd_kernel", "cagp_ks_blocksparse_bwd"),
         (r"ks_dense_fwd_kernel", "cagp_ks_dense_fwd"), (r"ks_dense_bwd_ls_kernel", "cagp_ks_dense_bwd_ls")]


def pairs_per_launch(cfg, entry):
    n = cfg["n"]
    if entry == "cagp_ks_blocksparse_fwd" and cfg.get("predict"):
        return float(cfg["m"]) * n
    return float(n) * n  # Gram passes: n^2 ordered pairs per launch (dense: m_local = n on one GPU)


def read(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    h = rows[hi]
    idc, kc, mc, vc = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    launches = defaultdict(dict)
    names = {}
    for r in rows[hi + 1:]:
        if len(r) <= vc:
            continue
        try:
            v = float(r[vc].replace(",", ""))
        except ValueError:
            continue
        launches[r[idc]][r[mc]] = v
        names[r[idc]] = r[kc]
    return launches, names


def main():
    out = {"generated_by": "tools/gen_kernel_metrics.py", "sources": [], "configs": {}}
    for path in sys.argv[1:]:
        cfg_name = re.search(r"ncu_([a-z0-9]+)\.csv$", path).group(1)
        cfg = B.CONFIGS[cfg_name]
        launches, names = read(path)
        per_entry = defaultdict(list)
        for lid, m in launches.items():
            nm = names[lid]
            for pat, entry in ENTRY:
                if re.search(r"\b" + pat + r"\b", nm):
                    per_entry[entry].append((nm, m))
                    break
        table = {}
        for entry, ls in per_entry.items():
            f64 = lambda m: sum(m.get(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum", 0.0) for o in ("dadd", "dfma", "dmul"))
            f32 = lambda m: sum(m.get(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum", 0.0) for o in ("fadd", "ffma", "fmul"))
            is64 = cfg["dtype"] == "f64"
            cnt = (lambda m: max(f64(m), m.get("smsp__thread_inst_executed_pipe_fp64_pred_on.sum", 0.0))) if is64 else f32
            top = max(cnt(m) for _, m in ls)
            work = [(nm, m) for nm, m in ls if cnt(m) > 0.5 * top]  # launches that did the work (a deferred kernel returns at once)
            nm = work[0][0]
            avg = lambda key: sum(m.get(key, 0.0) for _, m in work) / len(work)
            pairs = pairs_per_launch(cfg, entry)
            pipe_cyc = avg("smsp__pipe_fp64_cycles_active.sum")
            thread_inst = sum((f64(m) if is64 else f32(m)) for _, m in work) / len(work)
            dmma = is64 and "dense" in entry
            # DMMA kernels: the d-op thread-instruction counters miss the tensor instructions and this capture has no
            # tensor-pipe counter, so no executed-instruction figure is derived for them (bench.py then reports frac null)
            fp_inst = None if dmma else thread_inst
            short = re.sub(r"\(.*$", "", nm).replace("void ", "").replace("cagp::", "")
            table[entry] = {
                "kernel": short, "launches_captured": len(work), "pairs_per_launch": pairs,
                "fp_inst_per_pair": None if fp_inst is None else fp_inst / pairs,
                "fp_inst_definition": "not derived: DMMA kernel (d-op counters exclude tensor instructions)" if dmma else
                                      ("DADD+DFMA+DMUL" if is64 else "FADD+FFMA+FMUL") + " thread instructions",
                "fp64_pipe_active_pct": 100.0 * pipe_cyc / max(avg("smsp__cycles_elapsed.sum"), 1.0) if is64 else None,
                "xu_thread_inst_per_pair": avg("smsp__thread_inst_executed_pipe_xu_pred_on.sum") / pairs,
                "warp_inst_per_pair_warp": avg("smsp__inst_executed.sum") * 32.0 / pairs,
                "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "registers": avg("launch__registers_per_thread"),
                "duration_ms_under_ncu": avg("gpu__time_duration.sum") / 1e6,
                "dram_bytes_per_launch": avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
                "dram_read_bytes": avg("dram__bytes_read.sum"), "dram_write_bytes": avg("dram__bytes_write.sum"),
                "source": f"profiles/{os.path.basename(path)}",
            }
        out["configs"][cfg_name] = table
        out["sources"].append(os.path.basename(path))
    dst = os.path.join(ROOT, "profiles", "kernel_metrics.json")
    old = {}
    if os.path.exists(dst):
        old = json.load(open(dst))
    merged = dict(old.get("configs", {}))
    merged.update(out["configs"])
    out["configs"] = merged
    out["sources"] = sorted(set(old.get("sources", [])) | set(out["sources"]))
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    for c, t in out["configs"].items():
        for e, v in t.items():
            print(f"{c:6s} {e:32s} {v['kernel'][:60]:60s} fp/pair {v['fp_inst_per_pair']}  pipe {v['fp64_pipe_active_pct']}  dram {v['dram_bytes_per_launch'] / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
