#!/usr/bin/env python
"""Summaries of the SMF ensemble captures (bench.py --workload c5):
  gpurun_out/launches_c5_<tag>.csv  -> profiles/<tag>_c5_launches.txt
  gpurun_out/prof_smf_<tag>.ncu-rep -> profiles/<tag>_k_smf_run.txt, traffic.json[k_smf_run*]
usage: python profiles/summarize_smf.py r1 [chain_steps_in_the_captured_launch]"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
chain_steps = float(sys.argv[2]) if len(sys.argv) > 2 else 8192 * 250.0

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


p = os.path.join(OUT, "launches_c5_%s.csv" % tag)
if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    with open(os.path.join(HERE, "%s_c5_launches.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline\n")
        f.write("# launch 1 of k_smf_run is the 20000-proposal burn-in, the others one batch of 250 proposals on 8192 chains\n")
        f.write("%-4s %-28s %14s\n" % ("id", "kernel", "duration_us"))
        for r in rows[1:]:
            f.write("%-4s %-28s %14.1f\n" % (r[0], r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", "")) / 1e3))

rep = os.path.join(OUT, "prof_smf_%s.ncu-rep" % tag)
if os.path.exists(rep):
    r = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, units, vals = r[0], r[1], r[2]
    d, u = dict(zip(h, vals)), dict(zip(h, units))
    lines = ["# ncu --set full --clock-control none --import-source on -k regex:k_smf_run -s 4 -c 1  python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline",
             "kernel: %s" % d.get("Kernel Name", "k_smf_run"),
             "one launch = 250 proposals on each of 8192 chains at equilibrium (~560 segments per chain)"]
    for k in KEYS:
        if k in d:
            lines.append("%-80s %18s %s" % (k, d[k], u.get(k, "")))
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def num(k):
        try:
            return float(d[k].replace(",", "")) * mult.get(u.get(k, ""), 1.0)
        except Exception:
            return 0.0
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    inst = num("smsp__inst_executed.sum")
    act = num("smsp__thread_inst_executed_per_inst_executed.ratio")
    lines += ["",
              "per chain-step: DRAM %.0f B, warp instructions %.0f, thread instructions %.0f, global load sectors %.0f, "
              "global store sectors %.0f, local load sectors %.0f" % (
                  dram / chain_steps, inst / chain_steps, inst * act / chain_steps,
                  num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / chain_steps,
                  num("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum") / chain_steps,
                  num("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") / chain_steps)]
    src = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"])
    tmp = "/tmp/src_smf.csv"
    open(tmp, "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(HERE, "srclines.py"), tmp, "25"], capture_output=True, text=True).stdout
    # ncu files the out-of-line chain functions under the including file's name while keeping
    # their line numbers in smf_chain.h: show the text of that line
    chain_src = open(os.path.join(os.path.dirname(HERE), "lite_b200", "csrc", "smf_chain.h")).read().split("\n")
    fixed = []
    for ln in hot.split("\n"):
        parts = ln.split()
        if parts and parts[0].startswith("smf_ensemble") and len(parts) > 1 and parts[1].isdigit():
            k = int(parts[1])
            text = chain_src[k - 1].strip()[:90] if 0 < k <= len(chain_src) else ""
            if len(parts) >= 9:
                ln = "%-14s %5s %7s %7s %7s %6s %6s %6s %7s  %s" % (("smf_chain.h",) + tuple(parts[1:9]) + (text,))
        fixed.append(ln)
    lines += ["", "# hottest source lines (warp-stall samples)", "\n".join(fixed)]
    open(os.path.join(HERE, "%s_k_smf_run.txt" % tag), "w").write("\n".join(lines))
    tj = os.path.join(HERE, "traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    t["k_smf_run"] = dram
    t["k_smf_run_bytes_per_chain_step"] = dram / chain_steps
    t["_k_smf_run_source"] = "ncu --set full capture tag %s: one launch of 250 proposals x 8192 chains" % tag
    json.dump(t, open(tj, "w"), indent=1)
print("done")
