"""Kernel timing probe for tuning (not the benchmark): ACS energy on a generated CRTT
tessellation; prints k_split / k_reduce / merges+flips device times.
LITE_B200_LIB selects the library build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lite_b200
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**7
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
t, tau = lite_b200.generate_crtt(cells, seed=5); soa = t.export()
ctx = lite_b200.Context(0)
ctx.tess_upload(soa)
acs = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
th = np.array([11 / np.pi, 12**4 * 0.05, 1.0, -np.log(12.0)])
ctx.energy_set(acs); ctx.merges_flips()
sp, rd, rd2, mf = [], [], [], []
for r in range(reps):
    ctx.merges_flips(); ctx.clear_splits()
    ctx.add_splits_philox(n, 0x5EED, r * n)
    lpl, g, H = ctx.pl_eval(th)
    a, b, m = ctx.last_kernel_ms(); sp.append(a); mf.append(m); rd.append(ctx.last_reduce_ms())
    ctx.pl_eval(th); rd2.append(ctx.last_reduce_ms())
print("%s cells=%d n=%d | k_split %.4f ms (%.3e cand/s) | k_reduce first %.4f ms (%.0f GB/s) again %.4f ms (%.0f GB/s) | mf %.4f ms | lpl %.9f | status %s" % (
    os.path.basename(os.environ.get("LITE_B200_LIB", "default")), soa.n_cells(), n, np.median(sp), n / (np.median(sp) * 1e-3),
    np.median(rd), n * 32 / (np.median(rd) * 1e-3) / 1e9, np.median(rd2), n * 32 / (np.median(rd2) * 1e-3) / 1e9,
    np.median(mf), lpl, ctx.fetch("STATUS")), flush=True)
