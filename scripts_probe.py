import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)),'tests'))
import numpy as np, lite_b200
from util import load_golden
soa, gold = load_golden("t100")
ctx = lite_b200.Context(0)
print("fp64 peak TFLOP/s", ctx.measure_fp64_peak(), ctx.measure_fp64_peak())
ctx.tess_upload(lite_b200.TessSoA(soa))
acs = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
for feats in (acs, ["minus_is_segment_internal"], ["is_point_inside_window", "edge_length", "face_number", "face_area_2","face_perimeter", "face_shape", "face_sum_of_angles", "min_angle","seg_number", "is_segment_internal", "minus_is_segment_internal", "segment_size_2"]):
    ctx.clear_splits(); ctx.energy_set(feats); ctx.merges_flips()
    th = np.zeros(len(feats))
    for n in (10**6, 10**7, 10**7):
        ctx.clear_splits()
        ctx.add_splits_philox(n, 0x5EED, 0)
        lpl, g, H = ctx.pl_eval(th)
        a, b, m = ctx.last_kernel_ms()
        print(len(feats), n, "split ms %.3f -> %.3e cand/s | eval ms %.3f -> %.1f GB/s | mf ms %.3f" % (a, n/(a*1e-3), b, n*8*len(feats)/(b*1e-3)/1e9, m), flush=True)
cs = ctx.lines_sample_clip(10**8, 1, 0, 0); a,_,_ = ctx.last_kernel_ms(); print("lines checksum", cs, "ms %.3f -> %.3e lines/s" % (a, 1e8/(a*1e-3)))
