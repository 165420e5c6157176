"""The general-face source the GPU kernels run for faces with holes or doubly-bounded edges
(lite_b200/csrc/gen_face.h: ray_exit_face over all boundary cycles with the twin tie-break,
the four cases of hpolygon_insert_edge / hpolygon_remove_edge, lib/ttessel.cpp:4515-4576,
:4659-4925), compiled for the host by tests/cpp/gen_host.cpp and compared with the exact
oracle on the reference's own holed domains (tests/check_lite.cpp:129-229) and, because the
general code must also be right on ordinary faces, on the golden tessellations."""
import os
import subprocess

import numpy as np
import pytest

import lite_oracle as lo
from lite_b200.capi import FEATURE_IDS
from test_oracle_golden import holed_polygons, holed_polygons_for_flips
from util import ATOL_STAT, RTOL, load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "gen_host")
pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="tests/cpp/gen_host not built (make)")

ORDER = ["vx", "vy", "he_src", "he_twin", "he_next", "he_face", "he_seg", "he_next_hf", "he_prev_hf", "he_dir",
         "he_length", "face_inside", "face_he_offset", "face_he_count", "face_he_list", "seg_he_start", "seg_he_end",
         "seg_n_edges", "seg_on_boundary", "non_blocking", "blocking"]
DT = dict(vx="f8", vy="f8", he_dir="u1", he_length="f8", face_inside="u1", seg_on_boundary="u1")


def run_general(tmp_path, soa, feats, he, x, ang):
    d, nC = len(feats), len(he)
    a = {k: np.ascontiguousarray(soa[k], dtype=DT.get(k, "i4")) for k in ORDER}
    hd = np.array([len(a["vx"]), len(a["he_src"]), len(a["face_inside"]), len(a["face_he_list"]), len(a["seg_he_start"]),
                   len(a["non_blocking"]), len(a["blocking"]), d, nC], dtype=np.int32)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(hd.tobytes())
        for k in ORDER:
            f.write(a[k].tobytes())
        f.write(np.array([FEATURE_IDS[n] for n in feats], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(he, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(x, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ang, dtype=np.float64).tobytes())
    subprocess.run([EXE, fin, fout], check=True)
    raw = open(fout, "rb").read()
    pos = 0

    def take(dtype, n, shape=None):
        nonlocal pos
        v = np.frombuffer(raw, dtype=dtype, count=n, offset=pos)
        pos += v.nbytes
        return v.reshape(shape) if shape else v
    nNB, nFl = int(hd[5]), 2 * int(hd[6])
    out = dict(split_e2=take("i4", nC), split_p1=take("f8", 2 * nC, (nC, 2)), split_p2=take("f8", 2 * nC, (nC, 2)),
               split_stat=take("f8", nC * d, (nC, d)), merge_stat=take("f8", nNB * d, (nNB, d)),
               flip_e2=take("i4", nFl), flip_right=take("i4", nFl), flip_p2=take("f8", 2 * nFl, (nFl, 2)),
               flip_stat=take("f8", nFl * d, (nFl, d)))
    tail = [int(v) for v in take("i4", 10)]
    out["n_general"], out["overflow"] = tail[0], tail[1]
    out["insert_cases"], out["remove_cases"] = tail[2:6], tail[6:10]
    return out


def assert_stats(name, got, want, feats, area2_scale=0.0):
    """Relative 1e-9 with the usual absolute floor.  face_area_2 of a move is a difference of
    squared face areas as large as `area2_scale` (the reference's holed domains span 21 x 15
    units, so ~1e4): the oracle's own double sum carries a few ulp of THAT, which is the floor
    there."""
    assert got.shape == want.shape
    for k, f in enumerate(feats):
        g, w = got[:, k], want[:, k]
        tol = ATOL_STAT + RTOL * np.abs(w)
        if f == "face_area_2":
            tol = tol + 16 * np.finfo(float).eps * area2_scale
        bad = np.nonzero(np.abs(g - w) > tol)[0]
        assert len(bad) == 0, "%s feature %s: %d/%d differ, first %d: got %r want %r" % (
            name, f, len(bad), len(w), bad[0], g[bad[0]], w[bad[0]])


def check_against_oracle(tmp_path, t, n_splits, seed):
    soa = lo.export_soa(t)
    t.rnd = lo.Rng(seed)
    he, x, ang = lo.draw_split_candidates(t, n_splits)
    want = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), he, x, ang)
    got = run_general(tmp_path, soa, lo.ALL12, he, x, ang)
    assert got["overflow"] == 0
    np.testing.assert_array_equal(got["split_e2"], want["split_e2"])
    np.testing.assert_array_equal(got["flip_e2"], want["flip_e2"])
    np.testing.assert_array_equal(got["flip_right"], want["flip_right"])
    np.testing.assert_allclose(got["split_p1"], want["split_p1"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(got["split_p2"], want["split_p2"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(got["flip_p2"], want["flip_p2"], rtol=0, atol=1e-11)
    scale = max(lo.face_area_2(lo.face2poly(f), t) for f in t.internal_faces())
    assert_stats("split", got["split_stat"], want["split_stat"], lo.ALL12, scale)
    assert_stats("merge", got["merge_stat"], want["merge_stat"], lo.ALL12, scale)
    assert_stats("flip", got["flip_stat"], want["flip_stat"], lo.ALL12, scale)
    return got


def grow(window, seed, n_first_splits, n_moves):
    """check_predictions_4_random_smf (tests/check_lite.cpp:339-397): a few splits, then a mixed
    sequence of proposals, all applied, kept at a low segment count so that faces with holes
    and isthmus edges stay common."""
    t = lo.TTessel(lo.Rng(seed))
    t.insert_window(window)
    for i in range(n_first_splits + n_moves):
        if i < n_first_splits or i % 3 == 0:
            t.update(t.propose_split())
        elif i % 3 == 1 and t.non_blocking_segments:
            t.update(t.propose_merge())
        elif i % 3 == 2 and t.blocking_segments:
            t.update(t.propose_flip())
    return t


CASES = {"insert": np.zeros(4, dtype=int), "remove": np.zeros(4, dtype=int)}


@pytest.mark.parametrize("seed,n_moves", [(5, 0), (5, 12), (7, 30), (11, 60), (13, 90), (17, 18), (19, 45)])
def test_holed_domain_every_move_all12(tmp_path, seed, n_moves):
    t = grow(holed_polygons(), seed, 5, n_moves)
    got = check_against_oracle(tmp_path, t, 160, seed + 100)
    assert got["n_general"] >= 1
    CASES["insert"] += got["insert_cases"]
    CASES["remove"] += got["remove_cases"]


def _ccb_index(f, h):
    for i, c in enumerate([f.outer] + list(f.holes)):
        if any(e is h for e in lo.ccb(c)):
            return i
    raise KeyError


def _split_of_kind(t, kind, tries=4000):
    """A proposed split joining two points of one hole ("same_hole", cas 2 of
    hpolygon_insert_edge) or of two different holes ("two_holes", cas 4)."""
    for _ in range(tries):
        s = t.propose_split()
        f = s.e1.face
        if not f.holes:
            continue
        i1, i2 = _ccb_index(f, s.e1), _ccb_index(f, s.e2)
        if kind == "same_hole" and i1 == i2 and i1 > 0:
            return s
        if kind == "two_holes" and i1 > 0 and i2 > 0 and i1 != i2:
            return s
    raise RuntimeError("no %s split found" % kind)


@pytest.mark.parametrize("kinds", [("same_hole",), ("two_holes",), ("two_holes", "same_hole"), ("same_hole", "two_holes")])
def test_removal_of_hole_to_hole_segments(tmp_path, kinds):
    """Merges of segments that join a hole to itself or to another hole: cas 2 and cas 4 of
    hpolygon_remove_edge (lib/ttessel.cpp:4866-4922)."""
    t = lo.TTessel(lo.Rng(23))
    t.insert_window(holed_polygons())
    for k in kinds:
        t.update(_split_of_kind(t, k))
    got = check_against_oracle(tmp_path, t, 300, 31)
    CASES["insert"] += got["insert_cases"]
    CASES["remove"] += got["remove_cases"]


def test_every_case_of_insert_and_remove_was_met():
    """cas 1-4 of hpolygon_insert_edge and of hpolygon_remove_edge over the runs above."""
    assert (CASES["insert"] > 0).all() and (CASES["remove"] > 0).all(), CASES


@pytest.mark.parametrize("seed,n_moves", [(3, 6), (4, 25)])
def test_holed_domain_for_flips(tmp_path, seed, n_moves):
    t = grow(holed_polygons_for_flips(), seed, 4, n_moves)
    got = check_against_oracle(tmp_path, t, 120, seed)
    assert got["n_general"] >= 1


@pytest.mark.parametrize("name", ["t11", "t100"])
def test_general_code_on_ordinary_faces(tmp_path, name):
    """On hole-free simple faces the general path must give what the golden vectors hold."""
    soa, gold = load_golden(name)
    n = min(400, len(gold["cand_he"]))
    got = run_general(tmp_path, soa, lo.ALL12, gold["cand_he"][:n], gold["cand_x"][:n], gold["cand_angle"][:n])
    assert got["overflow"] == 0 and got["n_general"] == 0
    np.testing.assert_array_equal(got["split_e2"], gold["split_e2"][:n])
    np.testing.assert_array_equal(got["flip_e2"], gold["flip_e2"])
    np.testing.assert_array_equal(got["flip_right"], gold["flip_right"])
    assert_stats("split", got["split_stat"], gold["split_stat"][:n], lo.ALL12)
    assert_stats("merge", got["merge_stat"], gold["merge_stat"], lo.ALL12)
    assert_stats("flip", got["flip_stat"], gold["flip_stat"], lo.ALL12)


def test_general_code_on_the_committed_holed_golden(tmp_path):
    """tests/golden/holed.npz (the reference's holed domain, exact oracle) through the host build."""
    soa, gold = load_golden("holed")
    got = run_general(tmp_path, soa, lo.ALL12, gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    assert got["overflow"] == 0 and got["n_general"] >= 2
    np.testing.assert_array_equal(got["split_e2"], gold["split_e2"])
    np.testing.assert_array_equal(got["flip_e2"], gold["flip_e2"])
    scale = float(gold["area2_scale"])
    assert_stats("split", got["split_stat"], gold["split_stat"], lo.ALL12, scale)
    assert_stats("merge", got["merge_stat"], gold["merge_stat"], lo.ALL12, scale)
    assert_stats("flip", got["flip_stat"], gold["flip_stat"], lo.ALL12, scale)
