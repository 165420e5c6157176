"""Quick timing probe (not the benchmark): ACS energy on a generated CRTT tessellation."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, lite_b200
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**7
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.time(); t, tau = lite_b200.generate_crtt(cells, seed=5); soa = t.export()
print("generated", t.counts(), "in %.1fs" % (time.time() - t0), flush=True)
ctx = lite_b200.Context(0)
print("fp64 peak TFLOP/s", ctx.measure_fp64_peak())
t0 = time.time(); ctx.tess_upload(soa); print("upload ms", (time.time() - t0) * 1e3)
acs = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
th = np.array([11 / np.pi, 12**4 * 0.05, 1.0, -np.log(12.0)])
ctx.energy_set(acs); ctx.merges_flips()
for r in range(reps):
    ctx.clear_splits()
    ctx.add_splits_philox(n, 0x5EED, 0)
    lpl, g, H = ctx.pl_eval(th)
    a, b, m = ctx.last_kernel_ms()
    print("ACS n=%d split ms %.3f -> %.3e cand/s | eval ms %.3f -> %.1f GB/s | mf ms %.3f | lpl %.6f" % (n, a, n / (a * 1e-3), b, n * 32 / (b * 1e-3) / 1e9, m, lpl), flush=True)
print("status", ctx.fetch("STATUS"))
