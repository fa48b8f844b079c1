#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py <tag>      # reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_prof.ncu-rep

Writes profiles/<tag>_launches.md (per-kernel time shares of the bench command) and
profiles/<tag>_kernels.md (selected `--set full` counters per captured launch), and updates
profiles/ncu_traffic.json (dram bytes per launch per C-ABI entry point, read by bench.py).
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
HERE = os.path.join(ROOT, "profiles")

ENTRY_OF = {"k_hash_fwd": "b2n_hash_fwd", "k_hash_bwd_table": "b2n_hash_bwd",
            "itc::k_instant_fwd_tc": "b2n_instant_mlp_fwd_tc", "k_instant_fwd_tc": "b2n_instant_mlp_fwd_tc",
            "k_instant_bwd_tc": "b2n_instant_mlp_bwd_tc", "k_instant_fwd<": "b2n_instant_mlp_fwd",
            "k_instant_bwd<": "b2n_instant_mlp_bwd", "k_composite_fwd": "b2n_composite_fwd",
            "k_composite_bwd": "b2n_composite_bwd", "k_march_mask": "b2n_march_mask",
            "k_march_compact": "b2n_march_compact", "k_mlp256<0": "b2n_nerf_mlp_fwd", "k_mlp256<1": "b2n_nerf_mlp_bwd",
            "k_fmlp_fwd": "b2n_fmlp_fwd", "k_fmlp_bwd": "b2n_fmlp_bwd", "k_fmlp_wgrad": "b2n_fmlp_wgrad",
            "k_nerf_dx": "b2n_nerf_mlp_dx", "wg256::k_wgrad256": "b2n_nerf_mlp_wgrad", "k_wgrad256": "b2n_nerf_mlp_wgrad", "k_hash_bwd_input": "b2n_hash_bwd_input",
            "k_opt_prepare": "b2n_opt_prepare", "k_opt_adamw": "b2n_opt_adamw"}

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "smsp__inst_executed.sum",
           # L1TEX line visits ("wavefronts") and L2 sector traffic of global loads / reductions: the units the hash-grid
           # kernels are bound by (per (point, level): 4 gathers or reductions of an x-neighbour pair)
           "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_global_op_ld.sum" if False else "l1tex__data_pipe_lsu_wavefronts.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
           "lts__t_sectors_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_write.sum",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
           "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def short(name):
    name = name.split("(")[0]
    for pre in ("void ", "b2n::", "fm::", "m256::", "wg256::", "opt::", "itc::"):
        name = name.replace(pre, "")
    return name.strip()


def launches(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_val, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r[i_name])
        tot[k][0] += 1
        tot[k][1] += v
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(HERE, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n")
        f.write(f"All {sum(v[0] for v in tot.values())} launches of the command (5 steps + set-up), {total:.1f} ms total.\n\n")
        f.write("| kernel | launches | total ms | share | ours |\n|---|---:|---:|---:|---|\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
            ours = "yes" if k.startswith("k_") else "torch/ATen"
            f.write(f"| `{k[:70]}` | {n} | {ms:.3f} | {100 * ms / total:.1f} % | {ours} |\n")
        mine = sum(ms for k, (n, ms) in tot.items() if k.startswith("k_"))
        f.write(f"\nOur kernels: {100 * mine / total:.1f} % of the device time; torch glue (optimizer, TV loss, RNG, loss): "
                f"{100 * (1 - mine / total):.1f} %.\n")
    print("wrote", f"{tag}_launches.md")


def kernels(tag):
    import glob
    reps = sorted(glob.glob(os.path.join(OUT, f"{tag}_prof*.ncu-rep")))
    if not reps:
        return
    traffic_path = os.path.join(HERE, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    all_rows = []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        idx_ = {h: i for i, h in enumerate(rows[0])}
        for r in rows[2:]:
            all_rows.append((os.path.basename(rep), idx_, rows[1], r))
    with open(os.path.join(HERE, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of the bench commands, selected counters\n\n")
        for rep_name, idx, units, r in all_rows:
            name = short(r[idx["Kernel Name"]])
            f.write(f"## `{name}`  ({rep_name})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in idx and r[idx[m]] != "":
                    f.write(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |\n")
            f.write("\n")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd *= scale.get(units[idx["dram__bytes_read.sum"]], 1)
                wr *= scale.get(units[idx["dram__bytes_write.sum"]], 1)
                for k, entry in ENTRY_OF.items():
                    if name.startswith(k):
                        traffic[entry] = rd + wr
            except Exception:
                pass
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    # unit utilisation per C-ABI entry point (what actually bounds the kernel), read by bench.py into roofline.ncu
    units = {}
    want = {"dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "time_ms": "gpu__time_duration.sum"}
    for rep_name, idx, units_row, r in all_rows:
        name = short(r[idx["Kernel Name"]])
        for k, entry in ENTRY_OF.items():
            if name.startswith(k):
                d = {}
                for key, m in want.items():
                    if m in idx and r[idx[m]] != "":
                        v = float(r[idx[m]].replace(",", ""))
                        if key == "time_ms":
                            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units_row[idx[m]], 1e-6)
                        d[key] = round(v, 3)
                d["kernel"] = name[:60]
                d["capture"] = rep_name
                units.setdefault(entry, []).append(d)
    json.dump(units, open(os.path.join(HERE, "ncu_units.json"), "w"), indent=1, sort_keys=True)
    print("wrote", f"{tag}_kernels.md", "ncu_traffic.json and ncu_units.json")


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    kernels(tag)
