"""Shared helpers for the tests (golden loading, tolerances)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

# floating-point parity bar (BASELINE.json north_star): relative 1e-9 on the
# sums; per-candidate statistics get the same relative bar plus an absolute
# floor because the as-written edge_length delta is pure rounding noise
# (SURVEY.md §7 "Cancellation").
RTOL = 1e-9
ATOL_STAT = 1e-12


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    soa = {k[4:]: z[k] for k in z.files if k.startswith("soa_")}
    rest = {k: z[k] for k in z.files if not k.startswith("soa_")}
    return soa, rest


def inside_cum(soa):
    """cum / cum_he exactly as lite_tess_upload builds them: inside halfedges in
    slot order (faces in index order, each in ccb order from outer_ccb())."""
    lens = soa["he_length"]
    ids, cum = [], []
    c = 0.0
    for f in range(len(soa["face_inside"])):
        if not soa["face_inside"][f]:
            continue
        off, cnt = int(soa["face_he_offset"][f]), int(soa["face_he_count"][f])
        for h in soa["face_he_list"][off:off + cnt]:
            c += float(lens[h])
            cum.append(c)
            ids.append(int(h))
    return np.array(cum), np.array(ids, dtype=np.int32)
