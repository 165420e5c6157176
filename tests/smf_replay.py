"""Replay of an SMF chain trace (smf::TraceRec, lite_b200/csrc/smf_chain.h) in the
exact oracle: every proposal the chain made is rebuilt in oracle/lite_oracle.py
from its geometry (halfedge by endpoints, x, angle), and the oracle's crossed
edge, end points, Hastings ratio and energy variation must agree with the
chain's; accepted moves are applied to the oracle's tessellation, whose segment
counts and internal length must follow the chain's.  Used by the CPU test of the
chain source and by the GPU test of the ensemble kernel."""
import math

import numpy as np

import lite_oracle as lo

TRACE_DT = np.dtype([
    ("type", "f8"), ("status", "f8"), ("r", "f8"), ("e_var", "f8"), ("x_accept", "f8"),
    ("e1s", "f8", 2), ("e1t", "f8", 2), ("x", "f8"), ("angle", "f8"),
    ("p1", "f8", 2), ("p2", "f8", 2), ("e2s", "f8", 2), ("e2t", "f8", 2),
    ("n_seg", "f8"), ("n_nb", "f8"), ("n_b", "f8"), ("int_length", "f8"),
])
assert TRACE_DT.itemsize == 23 * 8

TOL_PT = 1e-9


def _close(p, q, tol=TOL_PT):
    return abs(float(p[0]) - q[0]) <= tol and abs(float(p[1]) - q[1]) <= tol


def _find_he(t, s, d):
    hits = [h for h in t.halfedges if _close(h.src.p, s) and _close(h.tgt.p, d)]
    assert len(hits) == 1, "halfedge (%r -> %r) found %d times" % (s, d, len(hits))
    return hits[0]


def replay(trace, names, thetas, p_split, p_merge, rect=(0, 0, 1, 1), rtol=1e-9):
    """Returns (oracle tessellation, number of evaluated proposals per type)."""
    t = lo.TTessel(lo.Rng(0))
    t.insert_window_rect(*rect)
    eng = lo.make_energy(names, thetas, t)
    chain = lo.SMFChain(eng, p_split, p_merge)
    scale = sum(abs(x) for x in thetas) + 1.0
    evaluated = [0, 0, 0]
    for i, rec in enumerate(trace):
        typ, status = int(rec["type"]), int(rec["status"])
        where = "step %d type %d" % (i, typ)
        if status == 0:
            # no candidate: the list the proposal draws from is empty (lib/ttessel.cpp:3281, :3293)
            assert typ in (1, 2), where
            assert len(t.non_blocking_segments if typ == 1 else t.blocking_segments) == 0, where
        else:
            assert status in (1, 2), where + ": the chain found no crossed edge"
            e1 = _find_he(t, rec["e1s"], rec["e1t"])
            if typ == 0:
                mod = lo.Split(e1, float(rec["x"]), float(rec["angle"]))
                assert _close(mod.p1, rec["p1"]) and _close(mod.p2, rec["p2"]), where
                assert _close(mod.e2.src.p, rec["e2s"]) and _close(mod.e2.tgt.p, rec["e2t"]), where + ": crossed edge"
                r, ev = chain.hasting_split(mod)
            elif typ == 1:
                assert e1.seg in t.non_blocking_segments, where
                mod = lo.Merge(e1)
                r, ev = chain.hasting_merge(mod)
            else:
                assert e1.seg in t.blocking_segments, where
                mod = lo.Flip(e1)
                assert _close(mod.p2, rec["p2"]), where
                assert _close(mod.e2.src.p, rec["e2s"]) and _close(mod.e2.tgt.p, rec["e2t"]), where + ": crossed edge"
                r, ev = chain.hasting_flip(mod)
            evaluated[typ] += 1
            assert abs(ev - rec["e_var"]) <= rtol * abs(ev) + 1e-12 * scale, (where, ev, rec["e_var"])
            assert abs(r - rec["r"]) <= 1e-8 * abs(r), (where, r, rec["r"])
            accept = r > 1 or rec["x_accept"] < r
            if abs(rec["x_accept"] - r) > 1e-8 * abs(r):
                assert accept == (status == 2), where + ": acceptance"
            if status == 2:
                t.update(mod)
                eng.value += ev
        n_int = len(t.all_segments) - t.number_of_window_edges()
        assert len(t.non_blocking_segments) == int(rec["n_nb"]), where
        assert len(t.blocking_segments) == int(rec["n_b"]), where
        assert n_int == int(rec["n_seg"]), where
        assert math.isclose(t.int_length, rec["int_length"], rel_tol=1e-11), where
    return t, evaluated
