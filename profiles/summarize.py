#!/usr/bin/env python
"""Turns the ncu captures in gpurun_out/ into the committed summaries of profiles/.

  python profiles/summarize.py r1

reads  gpurun_out/launches_<tag>.csv, prof_split_<tag>.ncu-rep, prof_reduce_<tag>.ncu-rep,
       prof_mf_<tag>.ncu-rep
writes profiles/<tag>_launches.txt   per-kernel launch list (count, total, share of the step)
       profiles/<tag>_<kernel>.txt   key ncu metrics + hottest source lines
       profiles/traffic.json         dram bytes per launch (read by bench.py -> roofline.traffic)
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches():
    p = os.path.join(OUT, "launches_%s.csv" % tag)
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    order = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0]
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += v
    step = {k: a for k, a in agg.items() if not k.startswith(("k_prep", "k_build", "k_dfma"))}
    tot = sum(a[1] for a in step.values())
    with open(os.path.join(HERE, "%s_launches.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 3 --warmup 3 --no-cpu-baseline\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes. share = of the kernels of a step\n")
        f.write("%-34s %6s %12s %12s %8s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k in order:
            a = agg[k]
            sh = ("%6.1f%%" % (100 * a[1] / tot)) if k in step else "  (setup)"
            f.write("%-34s %6d %12.1f %12.1f %8s\n" % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, sh))


def kernel(rep, name):
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return None
    raw = ncu(["-i", p, "--page", "raw", "--csv"])
    r = list(csv.reader(io.StringIO(raw)))
    h, units, vals = r[0], r[1], r[2]
    d = dict(zip(h, vals))
    u = dict(zip(h, units))
    lines = ["# ncu --set full --clock-control none --import-source on  (%s)" % rep, "kernel: %s" % d.get("Kernel Name", name)]
    for k in KEYS:
        if k in d:
            lines.append("%-72s %16s %s" % (k, d[k], u.get(k, "")))
    src = ncu(["-i", p, "--page", "source", "--print-source", "cuda,sass", "--csv"])
    tmp = os.path.join("/tmp", "src_%s.csv" % name)
    open(tmp, "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(HERE, "srclines.py"), tmp, "25"], capture_output=True, text=True).stdout
    lines.append("")
    lines.append("# hottest source lines (warp-stall samples; lite_b200/csrc/kernels.cuh)")
    lines.append(hot)
    open(os.path.join(HERE, "%s_%s.txt" % (tag, name)), "w").write("\n".join(lines))
    def num(k):
        try:
            return float(d[k].replace(",", ""))
        except Exception:
            return 0.0
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return num("dram__bytes_read.sum") * mult.get(u.get("dram__bytes_read.sum", "byte"), 1.0) + \
        num("dram__bytes_write.sum") * mult.get(u.get("dram__bytes_write.sum", "byte"), 1.0)


launches()
tj = os.path.join(HERE, "traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}
for rep, name in (("prof_split_%s.ncu-rep" % tag, "k_split"), ("prof_reduce_%s.ncu-rep" % tag, "k_reduce"),
                  ("prof_mf_%s.ncu-rep" % tag, "k_merges_flips"), ("prof_clip_%s.ncu-rep" % tag, "k_split_clip")):
    t = kernel(rep, name)
    if t is not None:
        traffic[name] = t
traffic["_source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, capture tag %s" % tag


def dp_flop(rep):
    """2*DFMA + DADD + DMUL thread instructions of one launch (what the kernel executes)."""
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return None
    r = list(csv.reader(io.StringIO(ncu(["-i", p, "--page", "raw", "--csv"]))))
    d = dict(zip(r[0], r[2]))

    def num(k):
        try:
            return float(d[k].replace(",", ""))
        except Exception:
            return 0.0
    per_cycle = (2 * num("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") +
                 num("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") +
                 num("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"))
    return per_cycle * num("smsp__cycles_elapsed.avg")


f = dp_flop("prof_split_%s.ncu-rep" % tag)
if f:
    traffic["k_split_dp_flop_per_candidate_executed"] = round(f / 1e7, 1)   # one launch = 10^7 candidates (C2)
    traffic["_dp_flop_source"] = "ncu %s: (2*dfma + dadd + dmul thread instructions per elapsed cycle) x elapsed cycles / candidates of the captured launch" % tag
f = dp_flop("prof_clip_%s.ncu-rep" % tag)
if f:
    traffic["k_split_clip_dp_flop_per_line_executed"] = round(f / 1.25e8, 1)  # one launch = 1.25e8 lines (C4)
json.dump(traffic, open(tj, "w"), indent=1)
print("wrote", sorted(os.listdir(HERE)))
