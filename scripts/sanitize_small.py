"""Small end-to-end pass over every kernel, for compute-sanitizer (memcheck):
  compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import math
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import lite_b200
from util import load_golden

soa, gold = load_golden("t100")
ALL12 = ["is_point_inside_window", "edge_length", "face_number", "face_area_2", "face_perimeter", "face_shape",
         "face_sum_of_angles", "min_angle", "seg_number", "is_segment_internal", "minus_is_segment_internal",
         "segment_size_2"]
ctx = lite_b200.Context(0)
for feats in (ALL12, ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]):
    ctx.tess_upload(lite_b200.TessSoA(soa))
    ctx.energy_set(feats)
    ctx.record_splits(True)
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    ctx.add_splits_philox(5000, 7, 0)
    print("pl", ctx.pl_eval(np.zeros(len(feats)))[0])
    ctx.clear_splits()
    ctx.record_splits(False)
    ctx.add_splits_philox(5001, 7, 0)
    print("pl", ctx.pl_eval(0.01 * np.ones(len(feats)))[0])
ctx.clear_splits()
print("lines", ctx.lines_sample_clip(7001, 3, 0, mode=0), ctx.lines_sample_clip(7001, 3, 0, mode=1))
ctx.smf_create(96, 120, seed=5)
ctx.smf_set_model(ALL12[:-1], [0.1, 0.7, -0.2, 3.0, 0.05, 0.02, 0.5, 0.4, -0.4, -1.2, 0.0], 0.33, 0.33)
rows = ctx.smf_run(2, 400)
print("smf all-features mean n", rows[-1, 1].mean())
ctx.smf_stats(validate=True)
tr = ctx.smf_trace(5, 50)
h = ctx.smf_export(3)
h.check()
ctx.smf_create(70, 40, seed=6)     # runs into the capacity limit on purpose
ctx.smf_set_model(["is_segment_internal"], [-math.log(6.0)], 0.33, 0.33)
try:
    ctx.smf_run(1, 3000, rows=False)
    print("no overflow?")
except lite_b200.LiteError as e:
    print("overflow reported:", str(e)[:60])
ctx.close()
print("done")
