"""The C++ host model (lite::HostTess, the state behind the TTessel facade)
replayed against the exact oracle on the same sequence of random splits,
merges and flips (the reference's squared_domain scenario,
tests/check_lite.cpp:446-542, :1031-1040)."""
import math
import os

import numpy as np
import pytest

import lite_b200
import lite_oracle as lo

pytestmark = pytest.mark.skipif(not os.path.exists(lite_b200.lib_path()), reason="libliteb200.so not built")


def _cells_from_soa(a):
    cells = []
    for f in np.nonzero(a["face_inside"])[0]:
        off, cnt = a["face_he_offset"][f], a["face_he_count"][f]
        hs = a["face_he_list"][off:off + cnt]
        pts = []
        for k, h in enumerate(hs):
            hp = hs[k - 1]
            # corner iff the previous halfedge does not continue straight (face2poly)
            if a["he_next_hf"][hp] < 0 or a["he_next_hf"][hp] != a["he_next"][hp]:
                v = a["he_src"][h]
                pts.append((round(float(a["vx"][v]), 10), round(float(a["vy"][v]), 10)))
        k = min(range(len(pts)), key=lambda i: pts[i])
        cells.append(tuple(pts[k:] + pts[:k]))
    return sorted(cells)


def _cells_from_oracle(t):
    cells = []
    for f in t.internal_faces():
        pts = [(round(float(p[0]), 10), round(float(p[1]), 10)) for p in lo.face2poly(f)[0]]
        k = min(range(len(pts)), key=lambda i: pts[i])
        cells.append(tuple(pts[k:] + pts[:k]))
    return sorted(cells)


def _find_he(a, p, q, tol=1e-11):
    sx, sy = a["vx"][a["he_src"]], a["vy"][a["he_src"]]
    tw = a["he_twin"]
    tx, ty = sx[tw], sy[tw]
    m = (np.abs(sx - float(p[0])) < tol) & (np.abs(sy - float(p[1])) < tol) & \
        (np.abs(tx - float(q[0])) < tol) & (np.abs(ty - float(q[1])) < tol)
    idx = np.nonzero(m)[0]
    assert len(idx) == 1
    return int(idx[0])


@pytest.mark.parametrize("seed", [5, 11])
def test_replay_against_oracle(seed):
    t = lo.TTessel(lo.Rng(seed))
    t.insert_window_rect(0, 0, 1, 1)
    h = lite_b200.HostTess((0, 0, 1, 1))
    n, m = 60, 120
    for i in range(n + m):
        a = h.export().a
        if i < n or i % 3 == 0:
            mod = t.propose_split()
            he = _find_he(a, mod.e1.src.p, mod.e1.tgt.p)
            # recover x, angle from the oracle's draw by re-deriving them is not
            # possible; feed the same numbers instead
            x, ang = mod._x, mod._angle
            h.update_split(he, x, ang)
        elif i % 3 == 1:
            if not t.non_blocking_segments:
                continue
            mod = t.propose_merge()
            he = _find_he(a, mod.e.src.p, mod.e.tgt.p)
            seg = a["he_seg"][he]
            h.update_merge(int(np.nonzero(a["non_blocking"] == seg)[0][0]))
        else:
            if not t.blocking_segments:
                continue
            mod = t.propose_flip()
            he = _find_he(a, mod.e1.src.p, mod.e1.tgt.p)
            seg = a["he_seg"][he]
            bi = int(np.nonzero(a["blocking"] == seg)[0][0])
            side = 1 if a["seg_he_end"][seg] == he else 0
            assert side == 1 or a["he_twin"][a["seg_he_start"][seg]] == he
            h.update_flip(bi, side)
        t.update(mod)
        h.check()
        cells, nnb, nb = h.counts()
        assert (cells, nnb, nb) == (len(t.internal_faces()), len(t.non_blocking_segments),
                                    len(t.blocking_segments))
        if i % 10 == 0 or i == n + m - 1:
            soa = h.export()
            assert _cells_from_soa(soa.a) == _cells_from_oracle(t)
            assert soa.int_length == pytest.approx(t.int_length, rel=1e-12)


def test_generator_reaches_equilibrium_estimate():
    """At equilibrium of the CRTT chain the closed-form PL estimate
    log(n_nb*pi/u) (reference demos/pseudo_lik_inference.cpp:80-81) is near log(tau)."""
    t, tau = lite_b200.generate_crtt(2000, seed=5)
    t.check()
    soa = t.export()
    cells, nnb, nb = t.counts()
    assert 0.7 * 2000 < cells < 1.4 * 2000
    est = math.log(nnb * math.pi / soa.int_length)
    assert abs(est - math.log(tau)) < 0.15
    # cells = n_s+1, halfedges = 6 n_s + 8
    ns = nnb + nb
    assert cells == ns + 1 and len(soa.a["he_src"]) == 6 * ns + 8


def test_text_roundtrip_and_fixture():
    here = os.path.dirname(os.path.abspath(__file__))
    text = open(os.path.join(here, "golden", "tessellation_data.txt")).read()
    h = lite_b200.HostTess(text=text)
    assert h.counts()[0] == 11
    h.check()
    h2 = lite_b200.HostTess(text=h.write_text())
    assert h2.counts() == h.counts()
    # the same cells as the exact oracle builds from the same file
    window, segs = lo.parse_text(text)
    assert _cells_from_soa(h.export().a) == _cells_from_oracle(lo.ttessel_from_segments(window, segs))


def test_reader_builds_cyclic_t_dependencies_at_once():
    """After flips a T-tessellation contains cycles of segments each ending on the next
    (no order of successive splits rebuilds them); the reference reads such a file by
    LineTes::insert_segment line by line (lib/ttessel.cpp:1377-1441, :904-980)."""
    h = lite_b200.HostTess((0, 0, 1, 1))
    h.crtt_run(4.0, 20000, 5)
    text = h.write_text()
    h2 = lite_b200.HostTess(text=text)
    h2.check()
    assert h2.counts() == h.counts()
    a, b = h.export(), h2.export()
    assert _cells_from_soa(a.a) == _cells_from_soa(b.a)
    assert b.int_length == pytest.approx(a.int_length, rel=1e-12)
    # the rebuilt model is a working one: the chain carries on from it
    h2.crtt_run(4.0, 3000, 1)
    h2.check()
    # the exact oracle reads the same file to the same cells (t100 is such a tessellation)
    t100 = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "t100.txt")).read()
    window, segs = lo.parse_text(t100)
    assert _cells_from_soa(lite_b200.HostTess(text=t100).export().a) == \
        _cells_from_oracle(lo.ttessel_from_segments(window, segs))


@pytest.mark.parametrize("bad,msg", [
    ("0/1 0/1 1/1 0/1 1/1 1/1 0/1 1/1\n1/4 0/1 1/4 1/2\n", "lies on no other segment"),       # end in the void
    ("0/1 0/1 1/1 0/1 1/1 1/1 0/1 1/1\n0/1 0/1 1/1 1/1\n", "L- or X-vertex"),                 # corner to corner
    ("0/1 0/1 1/1 0/1 1/1 1/1 0/1 1/1\n1/4 0/1 1/4 1/1\n0/1 1/2 1/1 1/2\n", "cross"),  # a crossing
    ("0/1 0/1 0/1 1/1 1/1 1/1 1/1 0/1\n", "orientation"),                                       # clockwise window
])
def test_reader_rejects_what_is_not_a_t_tessellation(bad, msg):
    with pytest.raises(lite_b200.LiteError, match=msg):
        lite_b200.HostTess(text=bad)


def test_host_model_on_a_non_convex_window():
    """An L-shaped window (read from the text format: one polygon line): the chain keeps a
    valid T-tessellation inside it, and the file round-trips."""
    ell = "0/1 0/1 4/1 0/1 4/1 2/1 2/1 2/1 2/1 4/1 0/1 4/1\n"
    h = lite_b200.HostTess(text=ell)
    assert h.counts() == (1, 0, 0)
    h.crtt_run(1.5, 5000, 3)
    h.check()
    cells, nnb, nb = h.counts()
    assert cells > 100 and cells == nnb + nb + 1
    a = h.export().a
    assert not np.any((a["vx"] > 2 + 1e-12) & (a["vy"] > 2 + 1e-12))     # nothing in the notch
    h2 = lite_b200.HostTess(text=h.write_text())
    h2.check()
    assert h2.counts() == h.counts()
    assert _cells_from_soa(h2.export().a) == _cells_from_soa(a)


def test_bench_fixture_is_the_generated_c2_tessellation():
    """bench.py's CPU arm loads tests/golden/c2_soa.npz (numpy only, no product library); the GPU
    arm generates its input with generate_crtt(10000, seed=5).  They must be the same tessellation."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_soa.npz"))
    t, _tau = lite_b200.generate_crtt(10000, seed=5)
    soa = t.export()
    assert float(z["int_length"]) == soa.int_length
    for k, v in soa.a.items():
        np.testing.assert_array_equal(z[k], v, err_msg=k)


@pytest.mark.parametrize("which,seed,n_moves", [("holed", 11, 60), ("holed", 5, 0), ("flips", 4, 25)])
def test_reader_takes_windows_with_holes(which, seed, n_moves):
    """LineTes::read's format with several polygons and holes (lib/ttessel.cpp:1395-1415): the
    host reader rebuilds the tessellation the exact oracle wrote, faces that surround a hole keep
    its boundary as an inner boundary (halfedges carrying the face's id, not listed on its outer
    boundary), the export matches the oracle's edge for edge, the text round-trips; the host
    model's moves refuse such a tessellation."""
    from test_gen_face_cpu import grow
    from test_oracle_golden import holed_polygons, holed_polygons_for_flips
    t = grow(holed_polygons() if which == "holed" else holed_polygons_for_flips(), seed, 5, n_moves)
    text = lo.write_text(t)
    h = lite_b200.HostTess(text=text)
    h.check()
    a, o = h.export().a, lo.export_soa(t)
    assert h.export().int_length == pytest.approx(o["int_length"], rel=1e-14)
    assert h.counts() == (int(np.sum(o["face_inside"])), len(o["non_blocking"]), len(o["blocking"]))

    def edges(s):
        """directed edge (end points) -> (bounds an inside face, listed on an outer boundary, edges of its segment)"""
        listed = np.zeros(len(s["he_src"]), dtype=bool)
        listed[s["face_he_list"]] = True
        out = {}
        for k in range(len(s["he_src"])):
            p, q = s["he_src"][k], s["he_src"][s["he_twin"][k]]
            key = (float(s["vx"][p]), float(s["vy"][p]), float(s["vx"][q]), float(s["vy"][q]))
            out[key] = (bool(s["face_inside"][s["he_face"][k]]), bool(listed[k]), int(s["seg_n_edges"][s["he_seg"][k]]),
                        bool(s["seg_on_boundary"][s["he_seg"][k]]))
        return out
    assert edges(a) == edges(o)
    inside = a["face_inside"][a["he_face"]].astype(bool)
    assert inside.sum() > len(a["face_he_list"])          # some halfedges lie on inner boundaries
    # text round trip
    h2 = lite_b200.HostTess(text=h.write_text())
    assert edges(h2.export().a) == edges(a)
    with pytest.raises(lite_b200.LiteError, match="holes"):
        h.update_merge(0)
    with pytest.raises(lite_b200.LiteError, match="holes"):
        h.crtt_run(3.0, 10, 1)
