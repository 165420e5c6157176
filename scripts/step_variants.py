"""Step-time variants: where the merge/flip kernels sit relative to the split kernel."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import lite_b200
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
t, _ = lite_b200.generate_crtt(10000, seed=5)
soa = t.export()
ctx = lite_b200.Context(0)
ctx.tess_upload(soa); ctx.energy_set(ACS)
th = np.array([3.5, 1036.8, 1.0, -2.48])
N = 10_000_000
def run(name, fn, reps=100):
    for k in range(10): fn(k)
    ctx.sync(); ctx.timer_mark(0)
    for k in range(reps): fn(k)
    ctx.timer_mark(1); ctx.sync()
    print("%-34s %.4f ms/step   split %.3f mf %.3f" % (name, ctx.timer_elapsed(0, 1) / reps, ctx.last_kernel_ms()[0], ctx.last_kernel_ms()[2]))
def v_mf_first(k):
    ctx.merges_flips(); ctx.clear_splits(); ctx.add_splits_philox(N, 7, k * N); ctx.pl_eval(th)
def v_split_first(k):
    ctx.clear_splits(); ctx.add_splits_philox(N, 7, k * N); ctx.merges_flips(); ctx.pl_eval(th)
def v_serial(k):
    ctx.merges_flips(); ctx.sync(); ctx.clear_splits(); ctx.add_splits_philox(N, 7, k * N); ctx.pl_eval(th)
def v_no_mf(k):
    ctx.clear_splits(); ctx.add_splits_philox(N, 7, k * N); ctx.pl_eval(th)
def v_eval_only(k):
    ctx.pl_eval(th)
ctx.merges_flips()
run("mf first", v_mf_first); run("split first", v_split_first); run("mf, sync, split", v_serial)
run("no mf (stale merge/flip stats)", v_no_mf); run("eval only", v_eval_only)
ctx.close()
