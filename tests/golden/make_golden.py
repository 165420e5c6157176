#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/ with the exact
oracle (oracle/lite_oracle.py).  Run from the repo root:

    python tests/golden/make_golden.py

Inputs: the reference's 11-cell fixture (tessellation_data.txt, copied from
reference wrap/R/RLiTe/inst/tessellation_data.txt) and T100, a CRTT chain as in
reference demos/pseudo_lik_inference.cpp:56-69 (SMFChain(&mod,0.33,0.33),
feature minus_is_segment_internal, theta=log(tau)) run by the oracle with
numpy PCG64 seed 5.  Outputs per tessellation: the SoA export, the exact text
form, a fed split list and the exact per-candidate answers for the 12 built-in
features, plus log-PL / gradient / Hessian for the CRTT and ACS energies.
"""
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import lite_oracle as lo  # noqa: E402


def build_t11():
    window, segs = lo.parse_text(open(os.path.join(HERE, "tessellation_data.txt")).read())
    return lo.ttessel_from_segments(window, segs, lo.Rng(5))


def build_t100(tau=2.6, steps=6000):
    t = lo.TTessel(lo.Rng(5))
    t.insert_window_rect(0, 0, 1, 1)
    e = lo.make_energy(["minus_is_segment_internal"], [math.log(tau)], t)
    lo.SMFChain(e, 0.33, 0.33).step(steps)
    return t


def emit(name, t, n_splits):
    t0 = time.time()
    soa = lo.export_soa(t)
    he, xs, angs = lo.draw_split_candidates(t, n_splits)
    e = lo.make_energy(lo.ALL12, None, t)
    res = lo.eval_candidates(t, e, he, xs, angs)
    out = {("soa_" + k): v for k, v in soa.items()}
    out.update(cand_he=he, cand_x=xs, cand_angle=angs)
    out.update(res)
    out["features"] = np.array(lo.ALL12)
    # PL values for two energies (column subsets of the 12-feature statistics)
    def cols(names):
        return [lo.ALL12.index(n) for n in names]
    for tag, names, thetas in (
            ("crtt", ["minus_is_segment_internal"], [[0.0], [1.3]]),
            ("acs", lo.ACS, [[0.0] * 4, lo.acs_theta()])):
        c = cols(names)
        for k, th in enumerate(thetas):
            v, g, H = lo.pl_from_stats(th, res["split_stat"][:, c], res["merge_stat"][:, c],
                                       res["flip_stat"][:, c], soa["int_length"])
            out["pl_%s_%d_theta" % (tag, k)] = np.array(th)
            out["pl_%s_%d_lpl" % (tag, k)] = np.float64(v)
            out["pl_%s_%d_grad" % (tag, k)] = g
            out["pl_%s_%d_hess" % (tag, k)] = H
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    open(os.path.join(HERE, name + ".txt"), "w").write(lo.write_text(t))
    print("%s: %d cells, %d nb, %d b, %d splits, %.1fs" % (
        name, len(t.internal_faces()), len(t.non_blocking_segments), len(t.blocking_segments),
        n_splits, time.time() - t0))


if __name__ == "__main__":
    t11 = build_t11()
    emit("t11", t11, 600)
    t100 = build_t100()
    n_nb, n_b = len(t100.non_blocking_segments), len(t100.blocking_segments)
    emit("t100", t100, 10000 - n_nb - 2 * n_b)
