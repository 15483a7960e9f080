#!/usr/bin/env python
"""Summarise ncu captures into profiles/ (tracked).  Usage:
    python tools/ncu_summary.py <tag> <gpurun_out/dir>      # e.g. r01a gpurun_out/r01a
Writes profiles/<tag>_launches.csv (copy of the launch list), profiles/<tag>_launches.md (per-kernel mean),
profiles/<tag>_<name>_raw.csv (ncu --page raw export of every .ncu-rep found) and profiles/<tag>_ncu.md (key metrics)."""
import collections
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_bytes.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    tag, d = sys.argv[1], sys.argv[2]
    os.makedirs("profiles", exist_ok=True)
    ll = os.path.join(d, "launches.csv")
    if os.path.exists(ll):
        shutil.copy(ll, f"profiles/{tag}_launches.csv")
        rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
        hdr = rows[0]
        agg = collections.OrderedDict()
        for r in rows[1:]:
            x = dict(zip(hdr, r))
            if x.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(x["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(x["Metric Unit"], 1.0)
            agg.setdefault((x["Kernel Name"], x["Grid Size"], x["Block Size"]), []).append(v)
        tot = sum(sum(v) for v in agg.values())
        with open(f"profiles/{tag}_launches.md", "w") as f:
            f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
            f.write("| launches | mean us | share of listed time | grid | block | kernel |\n|---|---|---|---|---|---|\n")
            for (k, g, b), v in agg.items():
                f.write(f"| {len(v)} | {sum(v)/len(v):.1f} | {sum(v)/tot:.3f} | {g} | {b} | `{k[:140]}` |\n")
    traffic = {}
    md = [f"# {tag}: ncu --set full captures (per launch)\n"]
    # .ncu-rep files brought back from the box, or — when they were too big to travel — their `--page raw --csv` export
    # made on the box (tools/gpu_r02_final.sh writes <name>_raw.csv and deletes the report)
    reps = sorted(glob.glob(os.path.join(d, "*.ncu-rep"))) + sorted(glob.glob(os.path.join(d, "prof_*_raw.csv")))
    for rep in reps:
        if rep.endswith(".csv"):
            name = os.path.basename(rep)[:-len("_raw.csv")]
            raw = open(rep).read()
        else:
            name = os.path.splitext(os.path.basename(rep))[0]
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(f"profiles/{tag}_{name}_raw.csv", "w").write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        md.append(f"\n## {name}\n")
        for r in rows[2:]:
            x = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            md.append(f"\n`{x['Kernel Name'][:160]}`\n\n| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in x:
                    md.append(f"| {k} | {x[k]} | {u[k]} |")
            try:
                rd = float(x["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
                wr = float(x["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
                md.append(f"| **dram traffic (read+write)** | {rd + wr:.0f} | byte |")
                traffic.setdefault(name, []).append({"kernel": x["Kernel Name"][:120], "dram_bytes": rd + wr,
                                                     "grid": x.get("launch__grid_size")})
            except (KeyError, ValueError):
                pass
    open(f"profiles/{tag}_ncu.md", "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(f"profiles/{tag}_traffic_detail.json", "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
