#!/usr/bin/env python
"""bench.py — pseudo-likelihood candidate moves evaluated per second.

Metric (BASELINE.json): PL candidate moves evaluated/sec at 1/2/4/8 B200 next to
the host-core CPU reference.  One candidate evaluated = sampled + clipped against
its face + statistic delta computed + accumulated once into log-PL / gradient /
Hessian.

A *step* is one pass of the hot path over one batch:
    SetEnergy (all merges + all flips)  ->  AddSplits(N_S) with the Philox
    sampler (sample + clip + delta)     ->  value, gradient and Hessian at theta
    (+ one NCCL allreduce of 1+d+d(d+1)/2 doubles when N > 1).

Workloads:
    default     : config C2 — ~10^4-cell CRTT tessellation, ACS energy d=4, 10^7
                  dummy splits PER GPU (weak scaling: N GPUs evaluate N*10^7 splits
                  of the same tessellation in contiguous shards).
    --workload c3 : config C3 — ~10^5-cell tessellation, 1.25*10^7 splits per GPU
                  (10^8 at N=8), the north star's multi-GPU sizing.
    --workload c4 : config C4 — loop 3a alone: sample isotropic lines (Philox) and clip them
                  against their faces on a ~10^5-cell tessellation, 1.25*10^8 lines per GPU per
                  step (10^9 at N=8), checksum mode (nothing stored); the store mode
                  (e1,e2,p1,p2,x,angle,U per line) is timed beside it.  Metric: lines/s.
    --workload c5 : config C5 — ensemble of independent SMF chains under the checkACS
                  energy (applications/checkACS.cpp, tau=6), 8192 chains per GPU (65536 at
                  N=8); a step is one checkACS batch of 250 proposals on every chain and the
                  metric is chain-steps/s.  No collective.

`value`  : whole-job candidates/s, tessellation already resident in HBM.
`e2e`    : the same through the C ABI starting from HOST buffers every step:
           lite_tess_upload (host SoA -> device) + energy + SetEnergy + AddSplits
           + eval, results read back to the host.
`--impl reference` : the CPU restatement of the reference's algorithm
           (oracle/oracle_cpu.cpp, all host threads) on a bounded sample of the
           same workload; the reference itself needs CGAL and cannot be built
           here.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints
# its version banner on stdout), so file descriptor 1 is pointed at stderr for the whole
# run and the result line goes to a private duplicate of the original stdout.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())

import numpy as np  # noqa: E402

ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
# demos/sim_acs.cpp:64-76 with tau=12, fi1=0.05, fi2=1, flattened in CatVector
# order edges, faces, faces, segs
THETA = [11.0 / math.pi, 12.0 ** 4 * 0.05, 1.0, -math.log(12.0)]
SEED = 0x5EED
# algorithmic work per split candidate
F_ALG_SPLIT = 1300.0   # DP flop: K1 sample+clip 300 + K2 delta 1000
B_ALG_SPLIT = 32.0     # bytes written: 8*d
B_ALG_REDUCE = 32.0    # bytes read:    8*d
F_ALG_REDUCE = 60.0


def workload(n_gpus, args):
    c3 = args.workload == "c3"
    cells = args.cells or (100000 if c3 else 10000)
    per_gpu = args.splits or (12_500_000 if c3 else 10_000_000)
    name = "custom" if (args.cells or args.splits) else ("C3" if c3 else "C2")
    return name, cells, per_gpu


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, period=None):
        if period is None:
            period = float(os.environ.get("LITE_BENCH_CLOCK_PERIOD", "0.003"))
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._halt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.sm)) if self.sm else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def soa_bytes(soa):
    return int(sum(a.nbytes for a in soa.a.values()) + 8)


F_ALG_CLIP = 300.0     # DP flop per line (sample + clip)
B_ALG_CLIP_STORE = 40.0  # bytes written per line in store mode (2 x i32 + 4 x f64)


def run_lines(args, rank, local_rank, world, nccl_id, dist, torch):
    """Workload C4: random-line sampling and face clipping throughput."""
    import lite_b200
    cells = args.cells or 100000
    per_gpu = args.splits or 125_000_000
    n_global = per_gpu * world
    t, _tau = lite_b200.generate_crtt(cells, seed=5)
    soa = t.export()
    ctx = lite_b200.Context(local_rank, rank, world, nccl_id)
    ctx.tess_upload(soa)
    fp64_peak = ctx.measure_fp64_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def allmax(v):
        if world > 1:
            tt = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return v

    for k in range(args.warmup):
        ctx.lines_sample_clip(n_global, SEED, k * n_global, mode=0)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    ms_k = 0.0
    ctx.timer_mark(0)
    for k in range(args.steps):
        cs = ctx.lines_sample_clip(n_global, SEED, (args.warmup + k) * n_global, mode=0)
        ms_k += ctx.last_kernel_ms()[0]
    ctx.timer_mark(1)
    barrier()
    ms_total = allmax(ctx.timer_elapsed(0, 1))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    value = n_global / (ms_total / args.steps * 1e-3)
    # end to end: tessellation from host buffers every step, checksum back to the host
    for k in range(2):
        ctx.tess_upload(soa)
        ctx.lines_sample_clip(n_global, SEED, 0, mode=0)
    barrier()
    ctx.timer_mark(2)
    for k in range(args.steps):
        ctx.tess_upload(soa)
        ctx.lines_sample_clip(n_global, SEED, k * n_global, mode=0)
    ctx.timer_mark(3)
    barrier()
    ms_e2e = allmax(ctx.timer_elapsed(2, 3))
    # store mode (64 B written per line, 40 B of them the reference's Split members)
    n_store = min(per_gpu, 50_000_000)
    ctx.clear_splits()
    ctx.lines_sample_clip(n_store * world, SEED, 0, mode=1)
    ms_store = 0.0
    for k in range(3):
        ctx.clear_splits()
        ctx.lines_sample_clip(n_store * world, SEED, k * n_store * world, mode=1)
        ms_store += ctx.last_kernel_ms()[0]
    ms_store /= 3
    status = ctx.fetch("STATUS")
    k_s = ms_k / args.steps * 1e-3
    rate = per_gpu / k_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    store_rate = n_store / (ms_store * 1e-3)
    roofline = {
        "kernel": "k_split<philox|checksum> (K1: Philox sample + clip, nothing stored)",
        "bound": "fp64", "achieved": rate * F_ALG_CLIP / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": rate * F_ALG_CLIP / 1e12 / fp64_peak,
        "peak_source": "DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 figure)",
        "alg_flop_per_line": F_ALG_CLIP, "lines_per_launch": per_gpu, "avg_launch_ms": k_s * 1e3, "traffic": None,
        "store_mode": {"lines_per_sec": store_rate, "avg_launch_ms": ms_store, "lines_per_launch": n_store,
                       "alg_bytes_per_line": B_ALG_CLIP_STORE, "written_bytes_per_line": 64,
                       "hbm_achieved_gbs": store_rate * B_ALG_CLIP_STORE / 1e9,
                       "hbm_frac": store_rate * B_ALG_CLIP_STORE / 1e9 / hbm_peak,
                       "hbm_frac_written": store_rate * 64 / 1e9 / hbm_peak},
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof))
            roofline["traffic"] = tr.get("k_split_clip")   # read-only: the L2-resident tessellation
            fx = tr.get("k_split_clip_dp_flop_per_line_executed")
            if fx:
                roofline["executed_flop_per_line"] = fx
                roofline["executed_frac"] = rate * fx / 1e12 / fp64_peak
        except Exception:
            pass
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_cpu as oc
        arrays = dict(soa.a)
        arrays["int_length"] = soa.int_length
        T = oc.Tess(arrays)
        cores = oc.hardware_threads()
        sample = args.cpu_sample or max(1_000_000, 1_000_000 * cores)
        oc.lines_clip(T, sample, SEED, 0, 0, sample // 10, cores)
        reps, t0 = 0, time.perf_counter()
        while True:
            oc.lines_clip(T, n_global, SEED, args.warmup * n_global, reps * sample, sample, cores)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 20:
                break
        dt = time.perf_counter() - t0
        cs_cpu, _bad = oc.lines_clip(T, 2_000_000, SEED, 7, 0, 2_000_000, cores)
        cs_gpu = ctx.lines_sample_clip(2_000_000 * world, SEED, 7, mode=0) if world == 1 else None
        cpu = {"value": sample * reps / dt, "unit": "lines/s", "cores": cores, "kind": "port",
               "sample": "%d passes of %d lines of the step's %d, %d threads; C++ double restatement of split_sample + "
                         "the Split constructor (oracle/oracle_cpu.cpp: oracle_lines_clip)" % (reps, sample, per_gpu, cores),
               "checksum_2e6_lines_equal": bool(cs_gpu is not None and int(cs_cpu[0]) == int(cs_gpu[0])
                                                and int(cs_cpu[1]) == int(cs_gpu[1]))}
    if rank == 0:
        line = {
            "metric": "random_lines_sampled_and_clipped_per_sec", "value": value, "unit": "lines/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: %d isotropic lines per GPU (%d in all) sampled (Philox, length-weighted "
                                   "halfedge, x uniform, angle = acos(1-2U)) and clipped against the faces of a %d-cell "
                                   "CRTT tessellation of the unit square; checksum mode" % (per_gpu, n_global, soa.n_cells()),
                       "cells": soa.n_cells(), "lines_per_gpu": per_gpu, "lines_global": n_global,
                       "l2": "the %0.1f MB tessellation is L2-resident by design and nothing is stored in checksum "
                             "mode; store mode writes 64 B per line (%d MB per launch, larger than L2); no flush"
                             % (soa_bytes(soa) / 1e6, n_store * 64 // 1_000_000),
                       "sharding": "contiguous shards of the batch, tessellation replicated, no collective",
                       "no_exit_lines": int(status[0])},
            "clocks": clocks,
            "e2e": {"value": n_global / (ms_e2e / args.steps * 1e-3), "unit": "lines/s",
                    "h2d_bytes_per_step": soa_bytes(soa), "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps,
                    "path": "lite_tess_upload(host SoA) + lite_lines_sample_clip -> host checksum"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "fp64_peak_tflops_measured": fp64_peak,
            "checksum": [int(cs[0]), int(cs[1])],
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    ctx.close()


def run_cpu_reference_lines(args, n_gpus, rank):
    """`--impl reference --workload c4`: split_sample + Split constructor on all host threads."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    cells = args.cells or 100000
    per_gpu = args.splits or 125_000_000
    arrays, _source = reference_tessellation(cells)
    T = oc.Tess(arrays)
    cores = oc.hardware_threads()
    sample = args.cpu_sample or max(1_000_000, 1_000_000 * cores)
    for _ in range(args.warmup):
        oc.lines_clip(T, sample, SEED, 0, 0, sample // 10, cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        oc.lines_clip(T, per_gpu * n_gpus, SEED, k * per_gpu * n_gpus, 0, sample, cores)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    emit({"impl": "reference", "metric": "random_lines_sampled_and_clipped_per_sec", "value": val, "unit": "lines/s",
          "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "C4: %d-cell CRTT tessellation; CPU arm clips a bounded sample of the step's lines"
                                 % (T.s.n_faces - 1), "cells": int(T.s.n_faces - 1), "lines_per_step_sampled": int(sample)},
          "cpu_baseline": {"value": val, "unit": "lines/s", "cores": cores, "kind": "port",
                           "sample": "%d of the step's %d lines per GPU, per step (oracle/oracle_cpu.cpp)" % (sample, per_gpu)},
          "e2e": {"value": val, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


SMF_TAU = 6.0                      # applications/checkACS.cpp:29
SMF_FEATURES = ["is_segment_internal", "edge_length"]
SMF_THETA = [-math.log(SMF_TAU), (SMF_TAU - 1) / math.pi]   # :113-117
SMF_NIPL = 250                     # proposals per batch (demos/sim_acs.cpp:21)
SMF_BURN = 20000                   # proposals before timing: the chain starts from the empty window
SMF_CAP = 1020                     # capacity in internal segments (equilibrium ~600 at tau=6)
# What one proposal has to move: the sampler's sums ((n_pairs/32 + 32) doubles,
# ~700 B at 600 segments), the rings of one or two faces (6 halfedges x 32 B each), the segment
# records and list slots (~64 B); an accepted move writes 2-3 new edge pairs, 1-2 vertices and
# relinks ~8 neighbours (~400 B).
B_ALG_SMF = 1500.0


def oracle_chain_rate(budget_s):
    """Exact-arithmetic oracle chain (oracle/lite_oracle.py SMFChain) from the empty
    window for about budget_s seconds: proposals per second on one core."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lite_oracle as lo
    t = lo.TTessel(lo.Rng(5))
    t.insert_window_rect(0, 0, 1, 1)
    eng = lo.make_energy(SMF_FEATURES, SMF_THETA, t)
    chain = lo.SMFChain(eng, 0.33, 0.33)
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        chain.step(25)
        done += 25
    dt = time.perf_counter() - t0
    return done / dt, done, len(t.all_segments) - 4


def run_smf(args, rank, local_rank, world, nccl_id, dist, torch):
    """Workload C5: ensemble of SMF chains; one step = SMF_NIPL proposals on every chain."""
    import lite_b200
    per_gpu = args.chains or 8192
    n_global = per_gpu * world
    ctx = lite_b200.Context(local_rank, rank, world, nccl_id)
    ctx.smf_create(n_global, SMF_CAP, seed=5)
    ctx.smf_set_model(SMF_FEATURES, SMF_THETA, 0.33, 0.33)
    n_local = ctx.smf_count()[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    ctx.smf_run(1, SMF_BURN, rows=False)          # untimed burn-in to the stationary size
    for _ in range(args.warmup):
        ctx.smf_run(1, SMF_NIPL, rows=False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    ms_kernel = 0.0
    ctx.timer_mark(0)
    for _ in range(args.steps):
        ctx.smf_run(1, SMF_NIPL, rows=False)
        ms_kernel += ctx.smf_last_ms()
    ctx.timer_mark(1)
    barrier()
    ms_total = ctx.timer_elapsed(0, 1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    # end to end: model set from host values and the batch's statistics rows read back, every step
    barrier()
    ctx.timer_mark(2)
    for _ in range(args.steps):
        ctx.smf_set_model(SMF_FEATURES, SMF_THETA, 0.33, 0.33)
        rows = ctx.smf_run(1, SMF_NIPL, rows=True)
    ctx.timer_mark(3)
    barrier()
    ms_e2e = ctx.timer_elapsed(2, 3)
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0].item()), float(tt[1].item())
    final = ctx.smf_stats(validate=True)
    per_step = n_global * SMF_NIPL
    value = per_step / (ms_total / args.steps * 1e-3)
    e2e_val = per_step / (ms_e2e / args.steps * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    k_s = ms_kernel / args.steps * 1e-3
    gbs = n_local * SMF_NIPL * B_ALG_SMF / k_s / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            per = json.load(open(prof)).get("k_smf_run_bytes_per_chain_step")
            if per:
                traffic = per * n_local * SMF_NIPL
        except Exception:
            pass
    roofline = {
        "kernel": "k_smf_run (one thread per chain: propose + Hastings ratio + update on a 16-bit DCEL in HBM)",
        "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
        "peak_source": hbm_src, "alg_bytes_per_chain_step": B_ALG_SMF,
        "chain_steps_per_launch": n_local * SMF_NIPL, "avg_launch_ms": k_s * 1e3, "traffic": traffic,
        "note": "latency bound (dependent loads of a pointer structure), not bandwidth bound: see profiles/",
    }
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, done, nseg = oracle_chain_rate(10.0)
        cpu = {"value": rate, "unit": "chain-steps/s", "cores": 1, "kind": "port",
               "sample": "the first %d proposals of ONE chain from the empty window (it reaches %d segments; the "
                         "timed GPU chains hold ~600), exact rational arithmetic as the CGAL reference "
                         "(oracle/lite_oracle.py), single thread as the reference" % (done, nseg)}
    if rank == 0:
        col = {n: i for i, n in enumerate(lite_b200.SMF_COLS)}
        line = {
            "metric": "smf_chain_steps_per_sec", "value": value, "unit": "chain-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "C5: %d independent SMF chains per GPU (%d in all) in the unit square under the checkACS "
                            "energy (is_segment_internal theta=-log 6, edge_length theta=5/pi, p_split=p_merge=0.33), "
                            "%d burn-in proposals from the empty window, then one step = one checkACS batch of %d "
                            "proposals on every chain" % (per_gpu, n_global, SMF_BURN, SMF_NIPL),
                "chains_per_gpu": per_gpu, "chains_global": n_global, "proposals_per_step_per_chain": SMF_NIPL,
                "mean_segments": float(final[col["n"]].mean()), "max_segments_capacity": SMF_CAP,
                "l2": "inputs larger than L2: the chains' tessellations take %d MB per GPU, touched at random; no flush"
                      % (ctx.smf_count()[3] * n_local // 1_000_000),
                "sharding": "chains [r*n/G, (r+1)*n/G) on rank r, Philox stream keyed by the global chain id, "
                            "no collective",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "chain-steps/s", "h2d_bytes_per_step": 12 * len(SMF_THETA) + 16,
                    "d2h_bytes_per_step": int(rows.nbytes), "ms_per_step": ms_e2e / args.steps,
                    "path": "lite_smf_set_model(host theta) + lite_smf_run -> host rows (12 columns per chain)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    ctx.close()


def run_cpu_reference_smf(args, n_gpus, rank):
    """`--impl reference --workload c5`: the exact oracle chain on one core, bounded sample per step."""
    if rank != 0:
        return
    rates = []
    # bounded: 2 s of chain per step, less if the whole run would pass two minutes
    budget = min(2.0, 120.0 / max(1, args.warmup + args.steps))
    for _ in range(args.warmup + args.steps):
        rate, done, nseg = oracle_chain_rate(budget)
        rates.append(rate)
    rates = rates[args.warmup:]
    val = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "smf_chain_steps_per_sec", "value": val, "unit": "chain-steps/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * budget,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5: one SMF chain under the checkACS energy from the empty window, exact arithmetic "
                               "(the reference is single-threaded; it needs CGAL and cannot be built here)",
                   "proposals_sampled_per_step": int(done)},
        "cpu_baseline": {"value": val, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                         "sample": "%.2f s of one chain from the empty window per step (%d proposals, %d segments)" % (budget, done, nseg)},
        "e2e": {"value": val, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


C2_FIXTURE = os.path.join(ROOT, "tests", "golden", "c2_soa.npz")


def reference_tessellation(cells):
    """The tessellation of the CPU arm WITHOUT loading the product library: the committed
    SoA export of the default C2 input (tests/golden/make_c2_fixture.py wrote it with the host
    generator, seed 5; tests/test_host_model_cpu.py checks it is what the GPU arm generates).
    Other sizes have no fixture and are generated through lite_b200's host model."""
    if cells == 10000 and os.path.exists(C2_FIXTURE):
        z = np.load(C2_FIXTURE)
        arrays = {k: z[k] for k in z.files}
        arrays["int_length"] = float(arrays["int_length"])
        return arrays, "tests/golden/c2_soa.npz (numpy only)"
    import lite_b200
    t, _tau = lite_b200.generate_crtt(cells, seed=5)
    soa = t.export()
    arrays = dict(soa.a)
    arrays["int_length"] = soa.int_length
    return arrays, "lite_b200.generate_crtt (no fixture for this size: the product's host generator builds the input)"


def run_cpu_reference(args, n_gpus, rank):
    """`--impl reference`: the reference's CPU algorithm (C++ double restatement,
    oracle/oracle_cpu.cpp) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    if args.workload == "c5":
        return run_cpu_reference_smf(args, n_gpus, rank)
    if args.workload == "c4":
        return run_cpu_reference_lines(args, n_gpus, rank)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    name, cells, per_gpu = workload(n_gpus, args)
    arrays, source = reference_tessellation(cells)
    T = oc.Tess(arrays)
    cores = oc.hardware_threads()
    n_mf = T.n_nb + 2 * T.n_b
    # bounded sample of the step's split batch: ~0.2 s of CPU work per step on 16 threads
    sample = args.cpu_sample or max(100_000, 60_000 * cores)
    for _ in range(args.warmup):
        oc.pl_pass(T, ACS, THETA, max(1000, sample // 10), SEED, 0, cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        v, g, H = oc.pl_pass(T, ACS, THETA, sample, SEED, k * sample, cores)
    dt = time.perf_counter() - t0
    per_step = sample + n_mf
    val = per_step * args.steps / dt
    line = {
        "impl": "reference",
        "metric": "pl_candidate_moves_evaluated_per_sec", "value": val, "unit": "candidates/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d-cell CRTT tessellation (unit square, host generator seed 5), ACS energy d=4; "
                               "CPU arm evaluates a bounded sample of the split batch" % (name, T.s.n_faces - 1),
                   "cells": int(T.s.n_faces - 1), "splits_per_step_sampled": int(sample),
                   "merges": int(T.n_nb), "flips": int(2 * T.n_b), "features": ACS, "tessellation_from": source},
        "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": "%d of the step's %d splits per GPU + all %d merges/flips, per step; C++ double "
                                   "restatement of the reference algorithm (reference needs CGAL: not buildable here)"
                                   % (sample, per_gpu, n_mf)},
        "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lpl": v,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--chains", type=int, default=0, help="c5: chains per GPU (default 8192)")
    ap.add_argument("--cells", type=int, default=0, help="override the tessellation size")
    ap.add_argument("--splits", type=int, default=0, help="override the dummy splits per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="splits per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short C3/C4/C5/Newton-fit runs of the default line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    if world > 1 and world != n_gpus:
        n_gpus = world

    if args.impl == "reference":
        run_cpu_reference(args, n_gpus, rank)
        return

    import torch
    import torch.distributed as dist
    import lite_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(lite_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    if args.workload in ("c4", "c5"):
        (run_smf if args.workload == "c5" else run_lines)(args, rank, local_rank, world, nccl_id, dist, torch)
        if world > 1:
            dist.destroy_process_group()
        return

    run_pl(args, rank, local_rank, world, nccl_id, dist, torch)
    if world > 1:
        dist.destroy_process_group()


ACS3 = ["face_area_2", "face_sum_of_angles", "is_segment_internal"]   # identifiable sub-vector


def newton_fit(ctx, n_splits, theta0, iters, stepsize=1.0, seed=SEED):
    """PLInferenceNOIS::Step (lib/ttessel.cpp:3470-3505) through the C ABI, store_path = true:
    per iteration gradient + Hessian at theta (ONE fused pass here, two host loops in the
    reference), theta += stepsize * (-H^-1 g), AddSplits(number of merges), GetValue at the new
    theta.  The dummy-split set starts at n_splits (config 2: 10^7) instead of the constructor's
    n_nb (:3458).  Returns the path and device / host times."""
    nm, _nf = ctx.merges_flips()
    ctx.clear_splits()
    ctx.reserve_splits(n_splits + (iters + 1) * max(nm, 1))   # the growth of the set never reallocates inside the fit
    theta = np.array(theta0, dtype=np.float64)
    path, values, gnorm = [theta.copy()], [], []
    h0 = time.perf_counter()
    ctx.timer_mark(4)
    ctx.add_splits_philox(n_splits, seed, 0)
    off = n_splits
    values.append(ctx.pl_eval(theta)[0])
    red_ms = 0.0
    evals = 1
    for _ in range(iters):
        _v, g, H = ctx.pl_eval(theta)
        red_ms += ctx.last_reduce_ms()
        theta = theta + stepsize * np.linalg.solve(H, -g)
        ctx.add_splits_philox(max(nm, 1), seed, off)
        off += max(nm, 1)
        v, g2, _ = ctx.pl_eval(theta)
        red_ms += ctx.last_reduce_ms()
        evals += 2
        path.append(theta.copy())
        values.append(v)
        gnorm.append(float(np.linalg.norm(g2)))
    ctx.timer_mark(5)
    dev_ms = ctx.timer_elapsed(4, 5)
    host_ms = (time.perf_counter() - h0) * 1e3
    return {"theta_path": [p.tolist() for p in path], "lpl_path": values, "grad_norm_after_step": gnorm,
            "device_ms": dev_ms, "host_ms": host_ms, "evaluations": evals, "k_reduce_ms_total": red_ms,
            "splits_final": ctx.count()[1], "merges": nm}


def run_secondary(ctx, lite_b200, soa_c2, fp64_peak, hbm_peak):
    """Short, driver-timed runs of the other BASELINE configs on this GPU (N=1): the Newton fit of
    config 2, one C3-sized shard, C4 (lines/s, checksum and store), C5 (chain-steps/s).  CUDA-event
    times on the context's stream; each entry carries its own roofline fraction."""
    out = {}
    fx = fxc = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        fx, fxc = tr.get("k_split_dp_flop_per_candidate_executed"), tr.get("k_split_clip_dp_flop_per_line_executed")
    except Exception:
        pass
    # ---- config 2 as BASELINE.json words it: the Newton fit ----------------------------
    try:
        ctx.clear_splits()
        ctx.energy_set(ACS3)
        n_fit = 10_000_000
        # start: the closed-form CRTT estimate for the segment parameter (the tessellation IS a CRTT
        # sample), zero for the two face parameters.  From theta = 0 the first plain
        # Newton step overshoots by six orders of magnitude in the face_area_2 direction (areas^2 ~ 1e-8).
        nm_c2 = ctx.merges_flips()[0]
        theta0 = [0.0, 0.0, -math.log(nm_c2 * math.pi / soa_c2.int_length)]
        newton_fit(ctx, 200_000, theta0, 1)             # warm-up: allocations, module load
        fit = newton_fit(ctx, n_fit, theta0, 4)
        per_eval_ms = fit["k_reduce_ms_total"] / (fit["evaluations"] - 1)
        gbs = n_fit * 8 * len(ACS3) / (per_eval_ms * 1e-3) / 1e9
        out["c2_newton_fit"] = {
            "what": "PLInferenceNOIS-equivalent fit of the ACS energy restricted to its identifiable sub-vector "
                    "(face_area_2, face_sum_of_angles, is_segment_internal; the as-written edge_length row of the "
                    "Hessian is singular, SURVEY.md App. C 3) on the C2 tessellation: 10^7 Philox dummy splits, then 4 "
                    "Newton iterations from (0, 0, closed-form CRTT estimate), each = gradient+Hessian pass, update, "
                    "AddSplits(n_merges), value pass",
            "theta_start": theta0,
            "time_to_fit_ms": fit["device_ms"], "host_ms": fit["host_ms"], "iterations": 4,
            "evaluations_over_all_splits": fit["evaluations"], "splits": fit["splits_final"],
            "theta_hat": fit["theta_path"][-1], "theta_path": fit["theta_path"], "lpl_path": fit["lpl_path"],
            "grad_norm_after_step": fit["grad_norm_after_step"],
            "candidate_evaluations_per_sec": (fit["splits_final"] * fit["evaluations"]) / (fit["device_ms"] * 1e-3),
            "roofline": {"kernel": "k_reduce<3> (one pass = value + gradient + Hessian)", "bound": "hbm",
                         "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "avg_launch_ms": per_eval_ms, "alg_bytes_per_candidate": 8 * len(ACS3),
                         "note": "passes alternate direction, so the end of the 240 MB block the previous pass "
                                 "touched last is still in the 126 MB L2: above 1 means L2 hits, not faster HBM"},
        }
        ctx.clear_splits()
        ctx.energy_set(ACS)
    except Exception as e:  # a secondary entry must never take the headline down
        out["c2_newton_fit"] = {"error": repr(e)}
    # ---- C3 / C4 on a 10^5-cell tessellation --------------------------------------------
    try:
        g0 = time.perf_counter()
        t3, _ = lite_b200.generate_crtt(100000, seed=5)
        soa3 = t3.export()
        gen_s = time.perf_counter() - g0
        c3 = lite_b200.Context(ctx_device(ctx))
        c3.tess_upload(soa3)
        c3.energy_set(ACS)
        nm3, nf3 = c3.merges_flips()
        theta = np.array(THETA)
        n3 = 12_500_000

        def step3(k):
            c3.merges_flips()
            c3.clear_splits()
            c3.add_splits_philox(n3, SEED, k * n3)
            return c3.pl_eval(theta)
        for k in range(3):
            step3(k)
        c3.sync()
        ks = kr = 0.0
        steps3 = 10
        c3.timer_mark(0)
        for k in range(steps3):
            step3(3 + k)
            ks += c3.last_kernel_ms()[0]
            kr += c3.last_reduce_ms()
        c3.timer_mark(1)
        ms = c3.timer_elapsed(0, 1) / steps3
        rate = n3 / (ks / steps3 * 1e-3)
        ent = {"what": "one GPU's shard of config 3: %d-cell tessellation, %d Philox splits + %d merges + %d flips per "
                       "step, ACS d=4" % (soa3.n_cells(), n3, nm3, nf3),
               "value": (n3 + nm3 + nf3) / (ms * 1e-3), "unit": "candidates/s", "ms_per_step": ms, "steps": steps3,
               "kernel_ms": {"k_split": ks / steps3, "k_reduce": kr / steps3}, "generator_s": gen_s,
               "roofline": {"kernel": "k_split", "bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac_nominal": rate * F_ALG_SPLIT / 1e12 / fp64_peak}}
        if fx:
            ent["roofline"]["achieved"] = rate * fx / 1e12
            ent["roofline"]["frac"] = rate * fx / 1e12 / fp64_peak
        out["c3_shard"] = ent
        # C4: sample + clip only
        n4 = 125_000_000
        for k in range(2):
            c3.lines_sample_clip(n4, SEED, k * n4, mode=0)
        kk = 0.0
        steps4 = 4
        c3.timer_mark(2)
        for k in range(steps4):
            cs = c3.lines_sample_clip(n4, SEED, (2 + k) * n4, mode=0)
            kk += c3.last_kernel_ms()[0]
        c3.timer_mark(3)
        ms4 = c3.timer_elapsed(2, 3) / steps4
        rate4 = n4 / (kk / steps4 * 1e-3)
        n_store = 50_000_000
        c3.clear_splits()
        c3.lines_sample_clip(n_store, SEED, 0, mode=1)
        st = 0.0
        for k in range(2):
            c3.clear_splits()
            c3.lines_sample_clip(n_store, SEED, k * n_store, mode=1)
            st += c3.last_kernel_ms()[0]
        st /= 2
        ent4 = {"what": "config 4: %d isotropic lines per step sampled and clipped against the faces of the %d-cell "
                        "tessellation, checksum mode; store mode (64 B per line) beside it" % (n4, soa3.n_cells()),
                "value": n4 / (ms4 * 1e-3), "unit": "lines/s", "ms_per_step": ms4, "steps": steps4,
                "kernel_ms": kk / steps4, "checksum": [int(cs[0]), int(cs[1])],
                "roofline": {"kernel": "k_split<philox|checksum>", "bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac_nominal": rate4 * F_ALG_CLIP / 1e12 / fp64_peak},
                "store_mode": {"lines_per_sec": n_store / (st * 1e-3), "kernel_ms": st, "bound": "hbm",
                               "hbm_frac_written": n_store * 64 / (st * 1e-3) / 1e9 / hbm_peak,
                               "hbm_frac_alg": n_store * B_ALG_CLIP_STORE / (st * 1e-3) / 1e9 / hbm_peak}}
        if fxc:
            ent4["roofline"]["achieved"] = rate4 * fxc / 1e12
            ent4["roofline"]["frac"] = rate4 * fxc / 1e12 / fp64_peak
        out["c4_lines"] = ent4
        c3.close()
    except Exception as e:
        out["c3_c4"] = {"error": repr(e)}
    # ---- C5: one GPU's 8192 chains ---------------------------------------------------------
    try:
        c5 = lite_b200.Context(ctx_device(ctx))
        n5 = 8192
        c5.smf_create(n5, SMF_CAP, seed=5)
        c5.smf_set_model(SMF_FEATURES, SMF_THETA, 0.33, 0.33)
        c5.smf_run(1, SMF_BURN, rows=False)
        for _ in range(2):
            c5.smf_run(1, SMF_NIPL, rows=False)
        c5.sync()
        steps5 = 8
        kk = 0.0
        c5.timer_mark(0)
        for _ in range(steps5):
            c5.smf_run(1, SMF_NIPL, rows=False)
            kk += c5.smf_last_ms()
        c5.timer_mark(1)
        ms5 = c5.timer_elapsed(0, 1) / steps5
        fin = c5.smf_stats(validate=True)
        col = {n: i for i, n in enumerate(lite_b200.SMF_COLS)}
        gbs = n5 * SMF_NIPL * B_ALG_SMF / (kk / steps5 * 1e-3) / 1e9
        out["c5_chains"] = {
            "what": "one GPU's share of config 5: %d SMF chains (checkACS energy), %d burn-in proposals, then one step = "
                    "%d proposals on every chain; every chain passes the is_valid check afterwards" % (n5, SMF_BURN, SMF_NIPL),
            "value": n5 * SMF_NIPL / (ms5 * 1e-3), "unit": "chain-steps/s", "ms_per_step": ms5, "steps": steps5,
            "kernel_ms": kk / steps5, "mean_segments": float(fin[col["n"]].mean()),
            "roofline": {"kernel": "k_smf_run", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "alg_bytes_per_chain_step": B_ALG_SMF,
                         "note": "latency bound (dependent loads of a pointer structure), see profiles/"}}
        c5.close()
    except Exception as e:
        out["c5_chains"] = {"error": repr(e)}
    return out


def ctx_device(ctx):
    return getattr(ctx, "device", 0)


def rel_gap(got, want):
    """max |got - want| / max |want| over an array: the scale of the whole quantity, because single
    entries of the gradient / Hessian (the as-written edge_length row) are pure rounding noise."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = float(np.max(np.abs(want))) if want.size else 0.0
    return float(np.max(np.abs(got - want))) / scale if scale > 0 else float(np.max(np.abs(got - want)))


def run_pl(args, rank, local_rank, world, nccl_id, dist, torch):
    """Workloads C2 / C3: the pseudo-likelihood evaluation pass."""
    import lite_b200
    name, cells, per_gpu = workload(world, args)
    t, _tau = lite_b200.generate_crtt(cells, seed=5)   # deterministic: identical on every rank
    soa = t.export()
    n_cells = soa.n_cells()
    n_global = per_gpu * world
    ctx = lite_b200.Context(local_rank, rank, world, nccl_id)
    ctx.tess_upload(soa)
    ctx.energy_set(ACS)
    nm, nf = ctx.merges_flips()
    fp64_peak = ctx.measure_fp64_peak()
    theta = np.array(THETA)
    per_step_global = n_global + nm + nf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def allmax(v):
        if world > 1:
            tt = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return v

    def step(k):
        ctx.merges_flips()
        ctx.clear_splits()
        ctx.add_splits_philox(n_global, SEED, k * n_global)
        return ctx.pl_eval(theta)

    def step_e2e(k):
        # everything from host buffers: tessellation SoA -> device, energy, candidates, results -> host
        ctx.tess_upload(soa)
        ctx.energy_set(ACS)
        ctx.merges_flips()
        ctx.clear_splits()
        ctx.add_splits_philox(n_global, SEED, k * n_global)
        return ctx.pl_eval(theta)

    # ---- device-resident arm -------------------------------------------------
    for k in range(args.warmup):
        step(k)
    # Clocks: one sampling thread in the whole job, on rank 0 (its GPU stands for the box), created
    # and started BEFORE the barrier.  nvmlInit takes milliseconds; in round 1 every rank ran it
    # between the barrier and its first timed step, so the ranks entered the loop up to ~7 ms apart
    # at N=8 and the first step's exchange absorbed that skew (0.37 ms per step over 20 steps: the
    # whole of the 0.54 weak-scaling efficiency the driver measured).
    which = os.environ.get("LITE_BENCH_SAMPLER", "rank0")
    sampler = ClockSampler(local_rank) if (which == "all" or (which == "rank0" and rank == 0)) else None
    if sampler:
        sampler.start()
    barrier()
    l0 = ctx.launch_count()
    host_ms = []
    ctx.timers_accumulated(reset=True)   # the library sums its kernel times: nothing is queried inside the loop
    ctx.timer_mark(0)
    for k in range(args.steps):
        h0 = time.perf_counter()
        lpl, g, H = step(args.warmup + k)
        host_ms.append((time.perf_counter() - h0) * 1e3)   # pl_eval returns after the stream has drained
    ctx.timer_mark(1)
    barrier()
    ms_split, ms_reduce, ms_mf, ms_exch, _n_eval = ctx.timers_accumulated()
    ms_total = ctx.timer_elapsed(0, 1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_total = allmax(ms_total)
    ms_step = ms_total / args.steps
    value = per_step_global / (ms_step * 1e-3)
    comm_mode = ctx.comm_info()[0]
    fails = ctx.pl_failures()

    # ---- end-to-end arm (host buffers every step) -------------------------------
    for k in range(3):
        step_e2e(k)
    sampler2 = ClockSampler(local_rank) if sampler else None   # the same GPU, over the second timed region
    if sampler2:
        sampler2.start()
    barrier()
    ctx.timer_mark(2)
    for k in range(args.steps):
        step_e2e(args.warmup + k)
    ctx.timer_mark(3)
    barrier()
    if sampler2 and clocks is not None:
        clocks["e2e_region"] = sampler2.stop()
    ms_e2e = allmax(ctx.timer_elapsed(2, 3))
    e2e_val = per_step_global / (ms_e2e / args.steps * 1e-3)
    d = len(ACS)
    # the caller's arrays + the cumulative halfedge lengths lite_tess_upload derives from them on the
    # host and sends along (8 B per listed halfedge) + theta and the feature ids
    h2d = soa_bytes(soa) + 8 * int(soa.a["face_he_list"].size) + 8 * d + 4 * d
    # what lite_pl_eval reads back: the result block (two 128-word sums, status words, the
    # exchange's report, the validation word: 268 doubles, one copy) and the per-block column sums
    # of the merge/flip kernel (12 doubles per block of 128 candidates), finished on the host
    d2h = 268 * 8 + ((nm + 127) // 128 + (nf + 127) // 128) * 12 * 8

    # ---- N > 1: the sharded evaluation against the same batch on ONE rank ------------
    shard_check = None
    if world > 1:
        n_chk = 4_000_001
        ctx.merges_flips()
        ctx.clear_splits()
        ctx.add_splits_philox(n_chk, SEED, 77)
        sh = ctx.pl_eval(theta)
        solo = lite_b200.Context(local_rank)            # world size 1 on this rank's GPU
        solo.tess_upload(soa)
        solo.energy_set(ACS)
        solo.merges_flips()
        solo.add_splits_philox(n_chk, SEED, 77)
        so = solo.pl_eval(theta)
        solo.close()
        gap = max(rel_gap(sh[0], so[0]), rel_gap(sh[1], so[1]), rel_gap(sh[2], so[2]))
        n_loc = ctx.count()[0]
        tt = torch.tensor([gap, float(n_loc)], dtype=torch.float64, device="cuda")
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        # every rank must also hold the same bits
        bits = torch.tensor(np.concatenate([[sh[0]], sh[1], sh[2].ravel()]), dtype=torch.float64, device="cuda")
        lo, hi = bits.clone(), bits.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same_bits = bool(torch.equal(lo, hi))
        shard_check = {"ok": bool(float(mx[0]) <= 1e-11 and int(sm[1].item()) == n_chk and same_bits),
                       "max_rel": float(mx[0]), "tol": 1e-11, "splits": n_chk, "shards_sum_to": int(sm[1].item()),
                       "identical_on_all_ranks": same_bits, "lpl_sharded": sh[0], "lpl_one_rank": so[0],
                       "what": "one batch of %d Philox splits (+ all merges/flips) evaluated in %d shards through the "
                               "job's contexts, and whole by a world-size-1 context on every rank: log-PL, gradient and "
                               "Hessian, max |diff| / max |value| per quantity, worst rank" % (n_chk, world)}
        ctx.clear_splits()

    # ---- roofline of the dominant kernel (k_split = sample + clip + delta) -----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    split_s = ms_split / args.steps * 1e-3
    red_s = ms_reduce / args.steps * 1e-3
    rate_split = per_gpu / split_s
    tf = rate_split * F_ALG_SPLIT / 1e12
    roofline = {
        "kernel": "k_split (Philox sample + clip + statistic delta, fused K1+K2)",
        "bound": "fp64", "unit": "TFLOP/s", "peak": fp64_peak,
        "peak_source": "DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 figure)",
        "achieved_nominal": tf, "frac_nominal": tf / fp64_peak,
        "alg_flop_per_candidate": F_ALG_SPLIT, "alg_bytes_per_candidate": B_ALG_SPLIT,
        "candidates_per_launch": per_gpu, "avg_launch_ms": split_s * 1e3,
        "hbm_achieved_gbs": rate_split * B_ALG_SPLIT / 1e9,
        "hbm_frac": rate_split * B_ALG_SPLIT / 1e9 / hbm_peak,
        "traffic": None,
    }
    gbs = per_gpu * B_ALG_REDUCE / red_s / 1e9
    roofline_reduce = {
        "kernel": "k_reduce<4> (theta-reduce: value + gradient + Hessian in one pass, K3)",
        "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
        "peak_source": hbm_src, "alg_bytes_per_candidate": B_ALG_REDUCE,
        "candidates_per_launch": per_gpu, "avg_launch_ms": red_s * 1e3, "traffic": None,
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    fx = None
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof))
            roofline["traffic"] = tr.get("k_split")
            roofline_reduce["traffic"] = tr.get("k_reduce")
            fx = tr.get("k_split_dp_flop_per_candidate_executed")
            roofline["executed_flop_source"] = tr.get("_dp_flop_source")
        except Exception:
            pass
    # `frac` is the HARDWARE fraction: DP flop the kernel executes per candidate (counted by ncu on
    # this build of the kernel, profiles/) x the live rate / the measured DFMA peak.  The nominal
    # figure (SURVEY.md §8d's 1300 flop of the reference ALGORITHM, which this kernel undercuts with
    # closed forms) stays beside it as frac_nominal and can exceed 1.
    if fx:
        roofline["executed_flop_per_candidate"] = fx
        roofline["achieved"] = rate_split * fx / 1e12
        roofline["frac"] = rate_split * fx / 1e12 / fp64_peak
    else:
        roofline["achieved"] = tf
        roofline["frac"] = tf / fp64_peak
        roofline["note"] = "no ncu flop count for this build under profiles/: frac is the nominal one"

    # ---- CPU baseline + in-run parity of the benched kernel (rank 0, N=1 only) ------
    cpu = None
    parity_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_cpu as oc
        arrays = dict(soa.a)
        arrays["int_length"] = soa.int_length
        T = oc.Tess(arrays)
        cores = oc.hardware_threads()
        sample = args.cpu_sample or max(100_000, 60_000 * cores)
        oc.pl_pass(T, ACS, THETA, sample // 10, SEED, 0, cores)
        reps, t0 = 0, time.perf_counter()
        while True:
            v_cpu, g_cpu, H_cpu = oc.pl_pass(T, ACS, THETA, sample, SEED, reps * sample, cores)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 20:
                break
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        oc.pl_pass(T, ACS, THETA, sample // 8, SEED, 0, 1)
        dt1 = time.perf_counter() - t1
        cpu = {"value": (sample + nm + nf) * reps / dt, "unit": "candidates/s", "cores": cores, "kind": "port",
               "sample": "%d passes of %d splits (of the step's %d) + all %d merges/flips, %d threads; C++ double "
                         "restatement of the reference algorithm (oracle/oracle_cpu.cpp)" % (
                             reps, sample, per_gpu, nm + nf, cores),
               "single_thread_value": (sample // 8 + nm + nf) / dt1}
        # the kernel instantiation the timed loop runs (Philox sampling + statistics, no record, the
        # compile-time ACS program) on the batch the port evaluated last: same seed, offset, size
        off = (reps - 1) * sample
        ctx.merges_flips()
        ctx.clear_splits()
        ctx.add_splits_philox(sample, SEED, off)
        v_gpu, g_gpu, H_gpu = ctx.pl_eval(theta)
        gaps = {"lpl": rel_gap(v_gpu, v_cpu), "grad": rel_gap(g_gpu, g_cpu), "hess": rel_gap(H_gpu, H_cpu)}
        parity_check = {"ok": bool(max(gaps.values()) <= 1e-9), "tol": 1e-9, "max_rel": max(gaps.values()),
                        "rel": gaps, "splits": int(sample), "seed": SEED, "offset": int(off),
                        "lpl_gpu": v_gpu, "lpl_cpu_port": v_cpu, "no_exit": list(ctx.pl_failures()),
                        "what": "k_split<philox|stat, ProgACS> + k_merges_flips + k_reduce<4> against oracle_cpu.pl_pass "
                                "on the same tessellation, seed, offset and batch size: max |diff| / max |value|"}
        ctx.clear_splits()

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary and args.workload == "c2" and not (args.cells or args.splits):
        secondary = run_secondary(ctx, lite_b200, soa, fp64_peak, hbm_peak)

    # ---- per-rank step diagnostics (where an N-GPU step spends its time) -------
    hm = np.array(host_ms)
    mine = {"rank": rank, "host_ms_min": float(hm.min()), "host_ms_median": float(np.median(hm)),
            "host_ms_max": float(hm.max()), "k_split_ms": ms_split / args.steps, "k_reduce_ms": ms_reduce / args.steps,
            "exchange_ms": ms_exch / args.steps}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    if rank == 0:
        line = {
            "metric": "pl_candidate_moves_evaluated_per_sec", "value": value, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "%s: PL evaluation pass of the sim_acs (ACS, d=4) energy on a %d-cell CRTT tessellation of "
                            "the unit square (host generator seed 5): all %d merges + %d flips + %d Philox dummy splits "
                            "per GPU, then value/gradient/Hessian at theta" % (name, n_cells, nm, nf, per_gpu),
                "cells": n_cells, "splits_per_gpu": per_gpu, "splits_global": n_global, "merges": nm, "flips": nf,
                "features": ACS, "theta": THETA,
                "l2": "inputs larger than L2: the per-step statistic block is %d MB per GPU (written by k_split, read "
                      "by k_reduce); the %0.1f MB tessellation is L2-resident by design; no flush" % (
                          per_gpu * 32 // 1_000_000, soa_bytes(soa) / 1e6),
                "sharding": "contiguous shards of the split batch, tessellation replicated; the %d split sums are added "
                            "across the GPUs %s" % (
                                1 + d + d * (d + 1) // 2,
                                {0: "(single rank: nothing to add)", 1: "by one ncclAllReduce per evaluation",
                                 2: "inside k_reduce's last block over NVLink peer memory (no NCCL kernel on the path)"}[comm_mode]),
            },
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "candidates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "path": "lite_tess_upload(host SoA) + lite_energy_set + lite_candidates_merges_flips + "
                            "lite_candidates_add_splits_philox + lite_pl_eval -> host lpl/grad/hess"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_reduce": roofline_reduce,
            "kernel_ms_per_step": {"k_split": ms_split / args.steps, "k_reduce_splits": ms_reduce / args.steps,
                                   "merges_flips": ms_mf / args.steps, "exchange": ms_exch / args.steps},
            "step_ms": {"min": float(hm.min()), "median": float(np.median(hm)), "max": float(hm.max()),
                        "what": "host clock around each timed step on rank 0 (every step ends with a stream "
                                "synchronisation); ms_per_step is the CUDA-event time of the whole loop / steps"},
            "collective": {"mode": {0: "none", 1: "nccl_allreduce", 2: "peer_memory_in_k_reduce"}[comm_mode],
                           "exchange_ms_per_step": ms_exch / args.steps, "per_rank": per_rank},
            "no_exit_candidates": list(fails),
            "fp64_peak_tflops_measured": fp64_peak,
            "lpl": lpl,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity_check is not None:
            line["parity_check"] = parity_check
        if shard_check is not None:
            line["shard_check"] = shard_check
        if secondary is not None:
            line["secondary"] = secondary
        emit(line)
    ctx.close()


if __name__ == "__main__":
    main()
