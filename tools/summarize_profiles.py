#!/usr/bin/env python
"""Turn a gpurun_out/<tag>/ directory (tools/gpu_round.sh) into the tracked summaries under profiles/.

  python tools/summarize_profiles.py gpurun_out/r01a r01
writes profiles/<round>_launches.txt   (per-kernel totals of the ncu launch list of bench.py)
       profiles/<round>_<kernel>_ncu.txt (key sections of the `ncu --set full` capture)
       profiles/<round>_bench.json / _bench_reference.json (the bench lines of the same visit)
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

KEYS = ["Duration", "Throughput", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "L1/TEX Hit Rate", "L2 Hit Rate",
        "Issued Ipc Active", "Issue Slots Busy", "No Eligible", "Eligible Warps", "Block Limit", "Warp Cycles Per Issued", "Grid Size", "Block Size",
        "Shared Memory Config", "Dynamic Shared", "Static Shared", "Mem Busy", "Max Bandwidth", "Mem Pipes Busy", "Executed Ipc", "SM Busy"]
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
       "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
       "sm__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:110]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as o:
        o.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline\n")
        o.write("# (covers synthetic-data generation by torch + 1 warm-up step + 1 timed step; times are serialised/cold, compare SHARES)\n")
        o.write(f"# total {tot:.3f} ms over {sum(c for c, _ in agg.values())} launches\n")
        o.write(f"{'ms':>11} {'share':>7} {'count':>6}  kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"{t:11.3f} {100 * t / tot:6.2f}% {c:6d}  {k}\n")


def ncu_summary(rep, dst, title):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    with open(dst, "w") as o:
        o.write(f"# {title}\n# source: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n\n")
        for line in det.splitlines():
            s = line.strip()
            if s.startswith("void") or "Context 1" in s or s.startswith("Section:") or any(k in s for k in KEYS):
                if not s.startswith(("OPT", "INF", "WRN")):
                    o.write(line.rstrip()[:160] + "\n")
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) > 2:
            hdr, units = rows[0], rows[1]
            o.write("\n# raw counters per captured launch\n")
            for r in rows[2:]:
                o.write(f"## launch id {r[0]}: {r[hdr.index('Kernel Name')][:80]} grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}\n")
                for k in RAW:
                    if k in hdr:
                        i = hdr.index(k)
                        o.write(f"{k:75s} {r[i]:>22s} {units[i]}\n")


def main():
    src, rnd = sys.argv[1], sys.argv[2]
    os.makedirs("profiles", exist_ok=True)
    if os.path.exists(f"{src}/launches.csv"):
        launches(f"{src}/launches.csv", f"profiles/{rnd}_launches.txt")
    for rep in sorted(os.listdir(src)):
        if rep.endswith(".ncu-rep"):
            ncu_summary(f"{src}/{rep}", f"profiles/{rnd}_{rep[:-8]}_ncu.txt", rep)
    for a, b in (("bench.json", "bench.json"), ("bench_ref.json", "bench_reference.json"), ("pytest_gpu.log", "pytest_gpu.log")):
        if os.path.exists(f"{src}/{a}"):
            shutil.copy(f"{src}/{a}", f"profiles/{rnd}_{b}")


if __name__ == "__main__":
    main()
