"""Where the end-to-end step spends its time (host side): LITE_B200_TRACE=1 python scripts/e2e_trace.py [cells [splits]]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import lite_b200
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_splits = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
t, _ = lite_b200.generate_crtt(cells, seed=5)
soa = t.export()
ctx = lite_b200.Context(0)
th = np.array([3.5, 1036.8, 1.0, -2.48])
for k in range(6):
    t0 = time.perf_counter(); ctx.tess_upload(soa)
    t1 = time.perf_counter(); ctx.energy_set(ACS)
    t2 = time.perf_counter(); ctx.merges_flips()
    t3 = time.perf_counter(); ctx.clear_splits(); ctx.add_splits_philox(n_splits, 7, 0)
    t4 = time.perf_counter(); ctx.pl_eval(th)
    t5 = time.perf_counter()
    print("upload %.3f energy %.3f mf %.3f splits %.3f eval(+wait) %.3f total %.3f ms" % tuple(
        1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)))
ctx.close()
