#!/usr/bin/env python
"""Golden vectors for a window WITH HOLES (tests/golden/holed.npz, holed.txt), made by the exact
oracle: the reference's `holed_polygons()` domain (reference tests/check_lite.cpp:129-200: two
polygons, seven holes) after five splits and sixty mixed moves (seed 11), 400 split candidates drawn
by the oracle's split_sample, every merge and flip, the twelve built-in features, and the
pseudo-likelihood for a small theta.  Run from the repo root:  python tests/golden/make_golden_holed.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "..", "oracle"), os.path.join(HERE, ".."), os.path.join(HERE, "..", "..")]
import lite_oracle as lo  # noqa: E402
from test_gen_face_cpu import grow  # noqa: E402
from test_oracle_golden import holed_polygons  # noqa: E402

t = grow(holed_polygons(), 11, 5, 60)
soa = lo.export_soa(t)
t.rnd = lo.Rng(2027)
he, xs, angs = lo.draw_split_candidates(t, 400)
res = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), he, xs, angs)
out = {("soa_" + k): v for k, v in soa.items()}
out.update(cand_he=he, cand_x=xs, cand_angle=angs)
out.update(res)
out["features"] = np.array(lo.ALL12)
th = 0.01 * np.random.default_rng(11).standard_normal(len(lo.ALL12))
th[lo.ALL12.index("face_area_2")] *= 1e-4
v, g, H = lo.pl_from_stats(th, res["split_stat"], res["merge_stat"], res["flip_stat"], soa["int_length"])
out.update(pl_theta=th, pl_lpl=np.float64(v), pl_grad=g, pl_hess=H)
out["area2_scale"] = np.float64(max(lo.face_area_2(lo.face2poly(f), t) for f in t.internal_faces()))
np.savez_compressed(os.path.join(HERE, "holed.npz"), **out)
open(os.path.join(HERE, "holed.txt"), "w").write(lo.write_text(t))
print("holed: %d cells (%d with holes), %d nb, %d b" % (len(t.internal_faces()), sum(1 for f in t.internal_faces() if f.holes),
                                                         len(t.non_blocking_segments), len(t.blocking_segments)))
