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
    """cum / cum_he exactly as lite_tess_upload builds them."""
    inside = soa["face_inside"][soa["he_face"]].astype(bool)
    ids = np.nonzero(inside)[0].astype(np.int32)
    cum = np.zeros(len(ids))
    c = 0.0
    lens = soa["he_length"]
    for k, h in enumerate(ids):
        c += float(lens[h])
        cum[k] = c
    return cum, ids
