"""GPU parity tests: the CUDA path, called through the C ABI, against the exact
oracle's committed golden vectors (tests/golden/*.npz, made by make_golden.py)
and against the oracle run live on device-sampled candidates.

Bar: bit-exact for candidate counts, face-hit / crossed-edge indices; relative
1e-9 for the floating-point sums (absolute floor where a statistic is pure
rounding noise)."""
import math

import numpy as np
import pytest

from util import ATOL_STAT, RTOL, inside_cum, load_golden

pytestmark = pytest.mark.gpu

ALL12 = ["is_point_inside_window", "edge_length", "face_number", "face_area_2",
         "face_perimeter", "face_shape", "face_sum_of_angles", "min_angle",
         "seg_number", "is_segment_internal", "minus_is_segment_internal", "segment_size_2"]
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]


@pytest.fixture(scope="module")
def lite():
    import lite_b200
    return lite_b200


def _ctx(lite, soa, features, record=True):
    ctx = lite.Context(0)
    ctx.tess_upload(lite.TessSoA(soa))
    ctx.energy_set(features)
    if record:
        ctx.record_splits(True)
    return ctx


# features whose delta is a difference of O(1) terms (angles in radians, perimeter / sqrt(area))
TERM_SCALE_ONE = ("face_shape", "face_sum_of_angles", "min_angle")


def _assert_stats(name, got, want, feats, term_scale=False):
    """term_scale: the 1e-9 bar is taken relative to the O(1) terms the delta is made of, not
    to the (possibly tiny) delta itself.  Needed against the DOUBLE port at scale: both sides
    round the ray's exit point, which the reference keeps exact, and next to a sliver face
    that rounding alone moves face_shape by ~1e-11 (see test_c2_scale_against_cpu_port)."""
    assert got.shape == want.shape
    for k, f in enumerate(feats):
        g, w = got[:, k], want[:, k]
        tol = ATOL_STAT + RTOL * np.abs(w)
        if term_scale and f in TERM_SCALE_ONE:
            tol = tol + RTOL
        bad = np.nonzero(np.abs(g - w) > tol)[0]
        assert len(bad) == 0, "%s feature %s: %d/%d candidates differ, first %d: got %r want %r" % (
            name, f, len(bad), len(w), bad[0], g[bad[0]], w[bad[0]])


@pytest.mark.parametrize("name", ["t11", "t100"])
def test_golden_candidates_all12(lite, name):
    soa, gold = load_golden(name)
    ctx = _ctx(lite, soa, ALL12)
    nm, nf = ctx.merges_flips()
    assert nm == len(gold["merge_he"]) and nf == len(gold["flip_e1"])
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    assert ctx.count() == (len(gold["cand_he"]), len(gold["cand_he"]))
    # integers: bit exact
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E1"), gold["cand_he"])
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), gold["split_e2"])
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), gold["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E1"), gold["flip_e1"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), gold["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), gold["flip_right"])
    assert ctx.fetch("STATUS")[0] == 0 and ctx.fetch("STATUS")[1] == 0
    # points
    np.testing.assert_allclose(ctx.fetch("SPLIT_P1"), gold["split_p1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(ctx.fetch("SPLIT_P2"), gold["split_p2"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ctx.fetch("FLIP_P2"), gold["flip_p2"], rtol=0, atol=1e-12)
    # statistics
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), gold["split_stat"], ALL12)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), gold["merge_stat"], ALL12)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), gold["flip_stat"], ALL12)
    ctx.close()


@pytest.mark.parametrize("name", ["t11", "t100"])
def test_golden_candidates_without_min_angle(lite, name):
    """Without min_angle the merge and flip kernels take their closed-form path (areas,
    perimeters and angles from the parents' precomputed features); same golden columns."""
    soa, gold = load_golden(name)
    feats = [f for f in ALL12 if f != "min_angle"]
    cols = [ALL12.index(f) for f in feats]
    ctx = _ctx(lite, soa, feats)
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), gold["split_stat"][:, cols], feats)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), gold["merge_stat"][:, cols], feats)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), gold["flip_stat"][:, cols], feats)
    ctx.close()


@pytest.mark.parametrize("name", ["t11", "t100"])
@pytest.mark.parametrize("tag,feats", [("crtt", ["minus_is_segment_internal"]), ("acs", ACS)])
def test_golden_pl_value_gradient_hessian(lite, name, tag, feats):
    soa, gold = load_golden(name)
    ctx = _ctx(lite, soa, feats, record=False)
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    for k in (0, 1):
        th = gold["pl_%s_%d_theta" % (tag, k)]
        lpl, g, H = ctx.pl_eval(th)
        wl, wg, wH = (gold["pl_%s_%d_%s" % (tag, k, s)] for s in ("lpl", "grad", "hess"))
        assert lpl == pytest.approx(float(wl), rel=RTOL)
        # the edges component of the ACS energy is rounding noise: absolute floor
        scale = np.abs(wg).max()
        np.testing.assert_allclose(g, wg, rtol=RTOL, atol=1e-12 * max(1.0, scale))
        np.testing.assert_allclose(H, wH, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wH).max()))
    ctx.close()


def test_crtt_closed_form_any_sample(lite):
    """lpl = th*n_nb - (u/pi)e^th - 2 n_b whatever the dummy sample
    (reference demos/pseudo_lik_inference.cpp:80-81)."""
    soa, gold = load_golden("t100")
    ctx = _ctx(lite, soa, ["minus_is_segment_internal"], record=False)
    nm, nf = ctx.merges_flips()
    ctx.add_splits_philox(100000, 0x5EED, 0)
    u = float(soa["int_length"])
    for th in (-0.5, 0.0, 1.7):
        lpl, g, H = ctx.pl_eval([th])
        assert lpl == pytest.approx(th * nm - (u / math.pi) * math.exp(th) - nf, rel=1e-12)
        assert g[0] == pytest.approx(nm - (u / math.pi) * math.exp(th), rel=1e-12)
        assert H[0, 0] == pytest.approx(-(u / math.pi) * math.exp(th), rel=1e-12)
    ctx.close()


def test_philox_sampler_matches_numpy_restatement(lite):
    import philox_ref
    soa, _ = load_golden("t100")
    ctx = _ctx(lite, soa, ACS)
    n, seed, off = 20000, 0x5EED, 12345
    ctx.add_splits_philox(n, seed, off)
    cum, cum_he = inside_cum(soa)
    he, x, U = philox_ref.sample_splits(cum, cum_he, n, seed, off, 0, n)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E1"), he)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_X"), x)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_U"), U)
    ang = ctx.fetch("SPLIT_ANGLE")
    np.testing.assert_allclose(ang, np.arccos(1 - 2 * U), rtol=0, atol=1e-15)  # device acos: <= 2 ulp
    # halfedges come out in sampler (slot) order, as split_sample's come out in container order
    # (lib/ttessel.cpp:1975-1989)
    order = {int(h): k for k, h in enumerate(cum_he)}
    pos = np.array([order[int(h)] for h in ctx.fetch("SPLIT_E1")])
    assert np.all(np.diff(pos) >= 0)
    ctx.close()


def test_philox_candidates_against_live_oracle(lite):
    """Device-sampled candidates re-evaluated by the exact oracle on the exact
    tessellation (text fixture)."""
    import lite_oracle as lo
    import os
    from util import GOLD
    soa, _ = load_golden("t100")
    window, segs = lo.parse_text(open(os.path.join(GOLD, "t100.txt")).read())
    t = lo.ttessel_from_segments(window, segs)
    soa2 = lo.export_soa(t)   # the oracle's own numbering of the rebuilt tessellation
    ctx = _ctx(lite, soa2, ALL12)
    ctx.merges_flips()
    n = 1500
    ctx.add_splits_philox(n, 7, 0)
    he, x, ang = ctx.fetch("SPLIT_E1"), ctx.fetch("SPLIT_X"), ctx.fetch("SPLIT_ANGLE")
    e = lo.make_energy(lo.ALL12, None, t)
    res = lo.eval_candidates(t, e, he, x, ang)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), res["split_e2"])
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), res["split_stat"], ALL12)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), res["merge_stat"], ALL12)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), res["flip_stat"], ALL12)
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), res["flip_e2"])
    ctx.close()


def test_extension_edge_length_internal(lite):
    import lite_oracle as lo
    import os
    from util import GOLD
    window, segs = lo.parse_text(open(os.path.join(GOLD, "t11.txt")).read())
    t = lo.ttessel_from_segments(window, segs, lo.Rng(3))
    soa = lo.export_soa(t)
    feats = ["edge_length_internal", "face_number"]
    ctx = _ctx(lite, soa, feats)
    ctx.merges_flips()
    he, x, ang = lo.draw_split_candidates(t, 200)
    ctx.add_splits_given(he, x, ang)
    res = lo.eval_candidates(t, lo.make_energy(feats, None, t), he, x, ang)
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), res["split_stat"], feats)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), res["merge_stat"], feats)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), res["flip_stat"], feats)
    ctx.close()


def test_nois_growth_and_clear(lite):
    """AddSplits appends (PLInferenceNOIS::Step, lib/ttessel.cpp:3500); ClearSplits empties."""
    soa, gold = load_golden("t100")
    ctx = _ctx(lite, soa, ACS, record=False)
    ctx.merges_flips()
    th = gold["pl_acs_1_theta"]
    he, x, ang = gold["cand_he"], gold["cand_x"], gold["cand_angle"]
    # add in three uneven chunks == add at once
    for a, b in ((0, 10), (10, 5000), (5000, len(he))):
        ctx.add_splits_given(he[a:b], x[a:b], ang[a:b])
    lpl, g, H = ctx.pl_eval(th)
    assert lpl == pytest.approx(float(gold["pl_acs_1_lpl"]), rel=RTOL)
    np.testing.assert_allclose(H, gold["pl_acs_1_hess"], rtol=RTOL, atol=1e-12 * np.abs(gold["pl_acs_1_hess"]).max())
    ctx.clear_splits()
    assert ctx.count() == (0, 0)
    with pytest.raises(lite.LiteError):
        ctx.pl_eval(th)
    ctx.close()


def test_rejects_bad_input(lite):
    soa, gold = load_golden("t11")
    ctx = lite.Context(0)
    with pytest.raises(lite.LiteError):
        ctx.energy_set(["face_number", "is_point_inside_window"])  # not in CatVector order
    bad = dict(soa)
    bad["int_length"] = float(soa["int_length"]) + 1.0
    with pytest.raises(lite.LiteError):
        ctx.tess_upload(lite.TessSoA(bad))
    ctx.tess_upload(lite.TessSoA(soa))
    ctx.energy_set(ACS)
    outside = int(np.nonzero(soa["face_inside"][soa["he_face"]] == 0)[0][0])
    with pytest.raises(lite.LiteError):
        ctx.add_splits_given([outside], [0.5], [1.0])
    ctx.close()


def test_random_lines_sample_clip_modes_and_oracle(lite):
    """Config "random_lines sampling and face clipping" (loop 3a alone): store mode
    against the CPU port of the oracle on the sampled (halfedge, x, angle), and the
    checksum mode against the checksum of what store mode wrote."""
    import oracle_cpu as oc
    soa, _ = load_golden("t100")
    ctx = _ctx(lite, soa, ACS, record=False)
    n, seed, off = 200000, 99, 5
    cs_store = ctx.lines_sample_clip(n, seed, off, mode=1)
    e1, e2 = ctx.fetch("SPLIT_E1", 0, n), ctx.fetch("SPLIT_E2", 0, n)
    p1, p2 = ctx.fetch("SPLIT_P1", 0, n), ctx.fetch("SPLIT_P2", 0, n)
    x, ang = ctx.fetch("SPLIT_X", 0, n), ctx.fetch("SPLIT_ANGLE", 0, n)
    cs_only = ctx.lines_sample_clip(n, seed, off, mode=0)
    np.testing.assert_array_equal(cs_store, cs_only)
    key = (e1.astype(np.uint64) << np.uint64(32)) | e2.astype(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        key = key * np.uint64(0x9E3779B97F4A7C15)
        key ^= key >> np.uint64(29)
    assert int(np.bitwise_xor.reduce(key)) == int(cs_only[0])
    assert int(e2.astype(np.uint32).astype(np.uint64).sum()) == int(cs_only[1])
    t = oc.Tess(soa)
    he, xx, aa = oc.sample_splits(t, n, seed, off, nthreads=4)
    np.testing.assert_array_equal(e1, he)
    np.testing.assert_array_equal(x, xx)
    r = oc.split_stats(t, ["face_number"], e1, x, ang, nthreads=4)
    assert r["bad"] == 0 and ctx.fetch("STATUS")[0] == 0
    np.testing.assert_array_equal(e2, r["split_e2"])
    np.testing.assert_allclose(p1, r["split_p1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(p2, r["split_p2"], rtol=0, atol=1e-12)
    # the CPU arm of `bench.py --workload c4` computes the same checksum
    cs_cpu, bad = oc.lines_clip(t, n, seed, off, nthreads=4)
    assert bad == 0
    np.testing.assert_array_equal(cs_only, cs_cpu)
    ctx.close()


def test_full_size_properties(lite):
    """BASELINE.json's full size (10^7 splits on a 10^4-cell tessellation) through
    size-independent properties: the CRTT value is the closed form whatever the sample;
    the per-feature column sums obey the conservation laws of the moves
    (is_segment_internal: +1 per split; face_area_2: -2*a1*a2 <= 0 and bounded by the
    parent; edge_length as written: rounding noise); evaluation is linear in the batch:
    two half batches appended == one batch."""
    t, _tau = lite.generate_crtt(10000, seed=5)
    soa = t.export()
    n = 10_000_000
    ctx = lite.Context(0)
    ctx.tess_upload(soa)
    ctx.energy_set(["minus_is_segment_internal"])
    nm, nf = ctx.merges_flips()
    ctx.add_splits_philox(n, 0x5EED, 0)
    u = soa.int_length
    for th in (-0.7, 2.0):
        lpl, g, H = ctx.pl_eval([th])
        assert lpl == pytest.approx(th * nm - (u / math.pi) * math.exp(th) - nf, rel=1e-12)
    assert ctx.fetch("STATUS")[0] == 0
    ctx.clear_splits()
    ctx.energy_set(ACS)
    ctx.merges_flips()
    th = np.array([11 / np.pi, 12.0 ** 4 * 0.05, 1.0, -np.log(12.0)])
    ctx.add_splits_philox(n, 7, 0)
    whole = ctx.pl_eval(th)
    # theta = 0: S0 = N, S1 = column sums of the stored statistics
    lpl0, g0, H0 = ctx.pl_eval(np.zeros(4))
    m, f = ctx.pl_sums()
    c = u / (math.pi * n)
    col = -(g0 + m + f + ctx.fetch("FLIP_STAT").sum(0)) / c      # sum over splits of -dphi
    assert col[3] == pytest.approx(-n, rel=1e-12)                 # is_segment_internal: +1 per split
    assert abs(col[0]) < 1e-6                                     # edge_length as written: noise
    assert 0 < col[1] < n * 0.5 * (1.0 / 100) ** 2 * 100           # 2*a1*a2, tiny cells
    assert col[2] < 0                                             # new corners only add to the sum
    ctx.clear_splits()
    ctx.add_splits_philox(n, 7, 0)          # same batch again: bitwise the same evaluation
    again = ctx.pl_eval(th)
    assert again[0] == whole[0]
    np.testing.assert_array_equal(again[2], whole[2])
    ctx.close()


def _assert_split_stats_adjudicated(a, got, want, feats, he, x, ang, max_outliers):
    """As _assert_stats, except that a handful of candidates may differ from the DOUBLE port
    in the geometric face features: those are recomputed the way the reference computes
    them (exact end points, tests/util.py:exact_split_face_stats) and the device must
    agree with that to the usual tolerance."""
    from util import exact_split_face_stats
    geometric = ("face_area_2", "face_perimeter", "face_shape", "face_sum_of_angles", "min_angle")
    n_out = 0
    for k, f in enumerate(feats):
        g, w = got[:, k], want[:, k]
        bad = np.nonzero(np.abs(g - w) > ATOL_STAT + RTOL * np.abs(w))[0]
        if len(bad) == 0:
            continue
        assert f in geometric, "split feature %s: %d candidates differ" % (f, len(bad))
        n_out += len(bad)
        assert n_out <= max_outliers, "split feature %s: %d candidates differ from the port" % (f, len(bad))
        for i in bad:
            exact = exact_split_face_stats(a, he[i], x[i], ang[i])[f]
            assert abs(g[i] - exact) <= ATOL_STAT + RTOL * abs(exact), (f, int(i), g[i], exact, w[i])
    return n_out


@pytest.mark.parametrize("feats_tag", ["acs", "all12"])
def test_c2_scale_against_cpu_port(lite, feats_tag):
    """Config C2's tessellation (10^4 cells, host generator seed 5) with 10^6 device-sampled
    splits and every merge and flip, against the C++ double port of the reference's
    algorithm (oracle/oracle_cpu.cpp, itself pinned to the exact oracle's golden vectors):
    halfedges, crossed edges, merge and flip identities bit exact; per-candidate
    statistics and log-PL / gradient / Hessian to relative 1e-9."""
    import oracle_cpu as oc
    feats = ACS if feats_tag == "acs" else ALL12
    t, _tau = lite.generate_crtt(10000, seed=5)
    soa = t.export()
    arrays = dict(soa.a)
    arrays["int_length"] = soa.int_length
    T = oc.Tess(arrays)
    n, seed, off = 1_000_000, 0x5EED, 3
    ctx = lite.Context(0)
    ctx.tess_upload(soa)
    ctx.energy_set(feats)
    ctx.record_splits(True)
    nm, nf = ctx.merges_flips()
    ctx.add_splits_philox(n, seed, off)
    he, x, ang = oc.sample_splits(T, n, seed, off, nthreads=8)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E1"), he)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_X"), x)
    r = oc.split_stats(T, feats, he, x, ctx.fetch("SPLIT_ANGLE"), nthreads=8)
    assert r["bad"] == 0 and ctx.fetch("STATUS")[0] == 0
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), r["split_e2"])
    np.testing.assert_allclose(ctx.fetch("SPLIT_P1"), r["split_p1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(ctx.fetch("SPLIT_P2"), r["split_p2"], rtol=0, atol=1e-12)
    # The port rounds p1 and p2 to doubles; the reference keeps them exact.  On the few
    # candidates whose chord is ~1e-7 long (or ends next to a vertex) that rounding moves an
    # acos-based feature by ~1e-10: there the exact construction decides (a few per 10^6).
    n_out = _assert_split_stats_adjudicated(soa.a, ctx.fetch("SPLIT_STAT"), r["split_stat"], feats, he, x,
                                            ctx.fetch("SPLIT_ANGLE"), max_outliers=200)
    assert n_out <= 200
    m = oc.merge_stats(T, feats, nthreads=8)
    f = oc.flip_stats(T, feats, nthreads=8)
    assert f["bad"] == 0 and nm == len(m["merge_he"]) and nf == len(f["flip_e1"])
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), m["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E1"), f["flip_e1"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), f["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), f["flip_right"])
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), m["merge_stat"], feats, term_scale=True)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), f["flip_stat"], feats, term_scale=True)
    rng = np.random.default_rng(1)
    for th in (np.zeros(len(feats)), 0.05 * rng.standard_normal(len(feats))):
        lpl, g, H = ctx.pl_eval(th)
        v, wg, wH = oc.pl(th, r["split_stat"], f["flip_stat"], m["merge_stat"].sum(0), f["flip_stat"].sum(0),
                          soa.int_length, nthreads=8)
        assert lpl == pytest.approx(v, rel=RTOL)
        np.testing.assert_allclose(g, wg, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wg).max()))
        np.testing.assert_allclose(H, wH, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wH).max()))
    ctx.close()


def test_convex_polygonal_window_against_live_oracle(lite):
    """Nothing on the path assumes a rectangle: a hexagonal window (TTessel::insert_window
    with a polygon, lib/ttessel.cpp:1516-1540), a tessellation grown by the exact oracle's
    SMF chain, then every merge and flip and 400 device-sampled splits for all twelve
    features against the exact oracle."""
    import lite_oracle as lo
    hexagon = [lo._pt(1, 0), lo._pt(3, 0), lo._pt(4, 2), lo._pt(3, 4), lo._pt(1, 4), lo._pt(0, 2)]
    t = lo.TTessel(lo.Rng(3))
    t.insert_window([(hexagon, [])])
    eng = lo.make_energy(["minus_is_segment_internal"], [math.log(1.2)], t)
    lo.SMFChain(eng, 0.4, 0.2).step(160)
    assert len(t.internal_faces()) > 12 and t.blocking_segments and t.non_blocking_segments
    soa = lo.export_soa(t)
    ctx = _ctx(lite, soa, ALL12)
    nm, nf = ctx.merges_flips()
    ctx.add_splits_philox(400, 11, 0)
    he, x, ang = ctx.fetch("SPLIT_E1"), ctx.fetch("SPLIT_X"), ctx.fetch("SPLIT_ANGLE")
    res = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), he, x, ang)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), res["split_e2"])
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), res["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E1"), res["flip_e1"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), res["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), res["flip_right"])
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), res["split_stat"], ALL12)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), res["merge_stat"], ALL12)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), res["flip_stat"], ALL12)
    # some candidates do start or end on the slanted window sides
    bnd = soa["seg_on_boundary"][soa["he_seg"]]
    assert bnd[he].any() and bnd[res["split_e2"]].any()
    ctx.close()


def test_non_convex_window_against_live_oracle(lite):
    """An L-shaped window (hole-free, non-convex): the cell at the reflex corner is not
    convex, so a line may cross its boundary more than twice and ray_exit_face's nearest-hit
    rule (lib/ttessel.cpp:4563) decides.  All twelve features, every move type, against the
    exact oracle."""
    import lite_oracle as lo
    ell = [lo._pt(0, 0), lo._pt(4, 0), lo._pt(4, 2), lo._pt(2, 2), lo._pt(2, 4), lo._pt(0, 4)]
    t = lo.TTessel(lo.Rng(3))
    t.insert_window([(ell, [])])
    eng = lo.make_energy(["minus_is_segment_internal"], [math.log(1.2)], t)
    lo.SMFChain(eng, 0.4, 0.2).step(120)
    soa = lo.export_soa(t)

    def convex(f):
        pts = [(float(p[0]), float(p[1])) for p in lo.face2poly(f)[0]]
        n = len(pts)
        return all((pts[(i + 1) % n][0] - pts[i][0]) * (pts[(i + 2) % n][1] - pts[(i + 1) % n][1]) -
                   (pts[(i + 1) % n][1] - pts[i][1]) * (pts[(i + 2) % n][0] - pts[(i + 1) % n][0]) > 0 for i in range(n))
    non_convex = [f for f in t.internal_faces() if not convex(f)]
    assert non_convex
    ctx = _ctx(lite, soa, ALL12)
    ctx.merges_flips()
    ctx.add_splits_philox(1500, 11, 0)
    he, x, ang = ctx.fetch("SPLIT_E1"), ctx.fetch("SPLIT_X"), ctx.fetch("SPLIT_ANGLE")
    in_non_convex = [i for i in range(len(he)) if any(t.halfedges[int(he[i])].face is f for f in non_convex)]
    assert len(in_non_convex) > 20
    res = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), he, x, ang)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), res["split_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), res["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), res["flip_right"])
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), res["split_stat"], ALL12)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), res["merge_stat"], ALL12)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), res["flip_stat"], ALL12)
    ctx.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_tessellations_against_live_oracle(lite, seed):
    """Six more tessellations grown by the exact oracle's chain from different seeds (with
    different proposal probabilities, so different mixes of blocking and non-blocking
    segments), each checked in full: every merge, every flip and 250 sampled splits, all
    twelve features, integers bit exact."""
    import lite_oracle as lo
    t = lo.TTessel(lo.Rng(100 + seed))
    t.insert_window_rect(0, 0, 1, 1)
    eng = lo.make_energy(["minus_is_segment_internal"], [math.log(1.0 + 0.5 * seed)], t)
    lo.SMFChain(eng, 0.25 + 0.05 * seed, 0.15).step(90)
    soa = lo.export_soa(t)
    ctx = _ctx(lite, soa, ALL12)
    nm, nf = ctx.merges_flips()
    assert nm == len(t.non_blocking_segments) and nf == 2 * len(t.blocking_segments)
    ctx.add_splits_philox(250, 1000 + seed, 0)
    he, x, ang = ctx.fetch("SPLIT_E1"), ctx.fetch("SPLIT_X"), ctx.fetch("SPLIT_ANGLE")
    res = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), he, x, ang)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), res["split_e2"])
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), res["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E1"), res["flip_e1"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), res["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), res["flip_right"])
    np.testing.assert_allclose(ctx.fetch("SPLIT_P1"), res["split_p1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(ctx.fetch("SPLIT_P2"), res["split_p2"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ctx.fetch("FLIP_P2"), res["flip_p2"], rtol=0, atol=1e-12)
    _assert_stats("split", ctx.fetch("SPLIT_STAT"), res["split_stat"], ALL12)
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), res["merge_stat"], ALL12)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), res["flip_stat"], ALL12)
    # and the pseudo-likelihood built from them
    th = 0.1 * np.random.default_rng(seed).standard_normal(len(ALL12))
    lpl, g, H = ctx.pl_eval(th)
    wl, wg, wH = lo.pl_from_stats(th, res["split_stat"], res["merge_stat"], res["flip_stat"], float(soa["int_length"]))
    assert lpl == pytest.approx(wl, rel=RTOL)
    np.testing.assert_allclose(g, wg, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wg).max()))
    np.testing.assert_allclose(H, wH, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wH).max()))
    ctx.close()


def test_host_generated_non_convex_window_against_cpu_port(lite):
    """A 350-cell tessellation of an L-shaped window grown by the host model's chain:
    20 000 sampled splits, every merge and flip, all twelve features against the C++ port."""
    import oracle_cpu as oc
    h = lite.HostTess(text="0/1 0/1 4/1 0/1 4/1 2/1 2/1 2/1 2/1 4/1 0/1 4/1\n")
    h.crtt_run(1.5, 5000, 3)
    soa = h.export()
    arrays = dict(soa.a)
    arrays["int_length"] = soa.int_length
    T = oc.Tess(arrays)
    ctx = lite.Context(0)
    ctx.tess_upload(soa)
    ctx.energy_set(ALL12)
    ctx.record_splits(True)
    ctx.merges_flips()
    n = 20000
    ctx.add_splits_philox(n, 5, 0)
    he, x, ang = oc.sample_splits(T, n, 5, 0, nthreads=4)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E1"), he)
    r = oc.split_stats(T, ALL12, he, x, ctx.fetch("SPLIT_ANGLE"), nthreads=4)
    assert r["bad"] == 0
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), r["split_e2"])
    _assert_split_stats_adjudicated(soa.a, ctx.fetch("SPLIT_STAT"), r["split_stat"], ALL12, he, x,
                                    ctx.fetch("SPLIT_ANGLE"), max_outliers=10)
    m, f = oc.merge_stats(T, ALL12, nthreads=4), oc.flip_stats(T, ALL12, nthreads=4)
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), m["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), f["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), f["flip_right"])
    _assert_stats("merge", ctx.fetch("MERGE_STAT"), m["merge_stat"], ALL12, term_scale=True)
    _assert_stats("flip", ctx.fetch("FLIP_STAT"), f["flip_stat"], ALL12, term_scale=True)
    ctx.close()


def test_benched_kernel_instantiation_is_value_checked(lite):
    """bench.py times k_split<philox|stat, ProgACS> with NO record.  (i) With the record
    on (the instantiation every integer parity test reads) the evaluation is bit-identical,
    for the compile-time ACS program and for a run-time program; (ii) the record-off
    evaluation agrees with the CPU port's pl_pass on the same seed / offset / batch."""
    import oracle_cpu as oc
    t, _tau = lite.generate_crtt(10000, seed=5)
    soa = t.export()
    n, seed, off = 1_000_001, 0x5EED, 12345
    th = np.array([11 / np.pi, 12.0 ** 4 * 0.05, 1.0, -np.log(12.0)])
    res = {}
    for feats, theta in ((ACS, th), (ACS[1:], th[1:])):
        for rec in (False, True):
            ctx = _ctx(lite, soa.a | {"int_length": soa.int_length}, feats, record=rec)
            ctx.merges_flips()
            ctx.add_splits_philox(n, seed, off)
            res[(len(feats), rec)] = ctx.pl_eval(theta)
            assert ctx.pl_failures() == (0, 0)
            ctx.close()
        a, b = res[(len(feats), False)], res[(len(feats), True)]
        assert a[0] == b[0]
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[2], b[2])
    arrays = dict(soa.a)
    arrays["int_length"] = soa.int_length
    v, g, H = oc.pl_pass(oc.Tess(arrays), ACS, th, n, seed, off, oc.hardware_threads())
    lpl, gg, HH = res[(4, False)]
    assert lpl == pytest.approx(v, rel=RTOL)
    np.testing.assert_allclose(gg, g, rtol=RTOL, atol=RTOL * np.abs(g).max())
    np.testing.assert_allclose(HH, H, rtol=RTOL, atol=RTOL * np.abs(H).max())


def test_failed_upload_leaves_no_tessellation(lite):
    """A rejected lite_tess_upload must not leave the previous tessellation half replaced:
    every later call that needs one fails with a state error instead of reading stale views."""
    soa, gold = load_golden("t100")
    ctx = _ctx(lite, soa, ACS, record=False)
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"][:100], gold["cand_x"][:100], gold["cand_angle"][:100])
    ctx.pl_eval(gold["pl_acs_1_theta"])
    small, _ = load_golden("t11")
    bad = dict(small)
    bad["int_length"] = float(small["int_length"]) + 1.0
    with pytest.raises(lite.LiteError):
        ctx.tess_upload(lite.TessSoA(bad))
    for call in (ctx.merges_flips, lambda: ctx.add_splits_given(gold["cand_he"][:4], gold["cand_x"][:4], gold["cand_angle"][:4]),
                 lambda: ctx.add_splits_philox(10, 1, 0), lambda: ctx.pl_eval(gold["pl_acs_1_theta"])):
        with pytest.raises(lite.LiteError, match="upload a tessellation"):
            call()
    assert ctx.count() == (0, 0)
    # and the context recovers with a good upload
    ctx.tess_upload(lite.TessSoA(soa))
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    assert ctx.pl_eval(gold["pl_acs_1_theta"])[0] == pytest.approx(float(gold["pl_acs_1_lpl"]), rel=RTOL)
    ctx.close()


def test_uploads_back_to_back_without_a_synchronisation(lite):
    """lite_tess_upload sends the block in groups from its packing threads, queues the checks and
    the preparation before the host has finished and sends the sampler's tables over the second
    stream.  None of that may depend on the caller synchronising: tessellations of different
    sizes and kinds (rectangle, window with holes -- which prepares twice --, rectangle) are
    uploaded one after the other with split kernels still queued in between, and the last
    evaluation of each equals a fresh context's, bit for bit."""
    t100, _ = load_golden("t100")
    t11, _ = load_golden("t11")
    holed, _ = load_golden("holed")
    import os
    th = np.array([0.3, -0.02, 0.1, -0.5])
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_soa.npz"))
    c2 = {k: z[k] for k in z.files}          # 10 810 cells: 3.8 MB per upload, split kernels of ~50 us
    c2["int_length"] = float(c2["int_length"])
    cases = ((t100, 50000, 1), (c2, 2000000, 5), (holed, 20000, 2), (t11, 30000, 3), (c2, 1000000, 6), (t100, 50000, 4))

    def fresh(a, n, seed):
        c = _ctx(lite, a, ACS, record=False)
        c.merges_flips()
        c.add_splits_philox(n, seed, 0)
        out = c.pl_eval(th)
        c.close()
        return out

    want = [fresh(a, n, seed) for a, n, seed in cases]
    ctx = lite.Context(0)
    got = []
    for rep in range(3):   # the second and third round reuse the arena and the slot arrays
        got = []
        for a, n, seed in cases:
            ctx.tess_upload(lite.TessSoA(a))
            ctx.energy_set(ACS)
            ctx.merges_flips()
            ctx.add_splits_philox(n, seed + 100, 0)      # left queued: the next lines do not wait for it
            ctx.clear_splits()
            ctx.tess_upload(lite.TessSoA(a))             # again, over the running kernels
            ctx.energy_set(ACS)
            ctx.merges_flips()
            ctx.add_splits_philox(n, seed, 0)
            got.append(ctx.pl_eval(th))
    for (v, g, H), (wv, wg, wH) in zip(got, want):
        assert v == wv
        np.testing.assert_array_equal(g, wg)
        np.testing.assert_array_equal(H, wH)
    ctx.close()


@pytest.mark.parametrize("which,seed,n_moves", [("holed", 5, 0), ("holed", 5, 12), ("holed", 11, 60), ("holed", 13, 90),
                                                 ("flips", 4, 25), ("hole_to_hole", 23, 0)])
def test_windows_with_holes_against_live_oracle(lite, which, seed, n_moves):
    """The reference's holed domains (tests/check_lite.cpp:129-229): faces with inner
    boundaries and isthmus edges, i.e. cas 2-4 of hpolygon_insert_edge / hpolygon_remove_edge and
    the twin tie-break of ray_exit_face (lib/ttessel.cpp:4554-4562, :4722-4792, :4866-4922),
    on the device (gen_launch.cuh beside the fast kernels).  Device-sampled splits (they land on
    inner boundaries too), every merge and every flip, all twelve features, integers bit exact,
    and the pseudo-likelihood built from them."""
    import lite_oracle as lo
    from test_gen_face_cpu import _split_of_kind, grow
    from test_oracle_golden import holed_polygons, holed_polygons_for_flips
    if which == "hole_to_hole":
        t = lo.TTessel(lo.Rng(seed))
        t.insert_window(holed_polygons())
        for k in ("two_holes", "same_hole"):
            t.update(_split_of_kind(t, k))
    else:
        t = grow(holed_polygons() if which == "holed" else holed_polygons_for_flips(), seed, 5, n_moves)
    soa = lo.export_soa(t)
    holed_faces = [f for f in t.internal_faces() if f.holes]
    assert holed_faces
    ctx = _ctx(lite, soa, ALL12)
    nm, nf = ctx.merges_flips()
    assert nm == len(t.non_blocking_segments) and nf == 2 * len(t.blocking_segments)
    ctx.add_splits_philox(400, 77 + seed, 0)
    he, x, ang = ctx.fetch("SPLIT_E1"), ctx.fetch("SPLIT_X"), ctx.fetch("SPLIT_ANGLE")
    on_holed = sum(1 for h in he if t.halfedges[int(h)].face.holes)
    on_inner = sum(1 for h in he if t.halfedges[int(h)].face.holes and
                   not any(e is t.halfedges[int(h)] for e in lo.ccb(t.halfedges[int(h)].face.outer)))
    assert on_holed > 20 and on_inner > 5
    res = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), he, x, ang)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), res["split_e2"])
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), res["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E1"), res["flip_e1"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), res["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), res["flip_right"])
    np.testing.assert_allclose(ctx.fetch("SPLIT_P1"), res["split_p1"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ctx.fetch("SPLIT_P2"), res["split_p2"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(ctx.fetch("FLIP_P2"), res["flip_p2"], rtol=0, atol=1e-11)
    # face_area_2 of a move is a difference of squared face areas of order `scale` (the domain
    # spans 21 x 15 units): a few ulp of that is the floor, for the oracle's own double sum too
    scale = max(lo.face_area_2(lo.face2poly(f), t) for f in t.internal_faces())
    k_a2 = ALL12.index("face_area_2")
    for name, got, want in (("split", ctx.fetch("SPLIT_STAT"), res["split_stat"]),
                            ("merge", ctx.fetch("MERGE_STAT"), res["merge_stat"]),
                            ("flip", ctx.fetch("FLIP_STAT"), res["flip_stat"])):
        assert np.all(np.abs(got[:, k_a2] - want[:, k_a2]) <= ATOL_STAT + RTOL * np.abs(want[:, k_a2]) + 16 * 2.3e-16 * scale), name
        keep = [k for k in range(len(ALL12)) if k != k_a2]
        _assert_stats(name, got[:, keep], want[:, keep], [ALL12[k] for k in keep])
    th = 0.01 * np.random.default_rng(seed).standard_normal(len(ALL12))
    th[k_a2] *= 1e-4
    lpl, g, H = ctx.pl_eval(th)
    assert ctx.pl_failures() == (0, 0)
    wl, wg, wH = lo.pl_from_stats(th, res["split_stat"], res["merge_stat"], res["flip_stat"], float(soa["int_length"]))
    assert lpl == pytest.approx(wl, rel=RTOL)
    np.testing.assert_allclose(g, wg, rtol=RTOL, atol=1e-9 * max(1.0, np.abs(wg).max()))
    np.testing.assert_allclose(H, wH, rtol=RTOL, atol=1e-9 * max(1.0, np.abs(wH).max()))
    # the fed list takes the same route
    ctx.clear_splits()
    ctx.add_splits_given(he, x, ang)
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), res["split_e2"])
    assert ctx.pl_eval(th)[0] == pytest.approx(wl, rel=RTOL)
    # and the clip-only mode counts them in its checksum
    cs = ctx.lines_sample_clip(400, 77 + seed, 0, mode=0)
    assert int(cs[1]) == int(res["split_e2"].astype(np.int64).sum())
    ctx.close()


def test_iid_sampler_is_the_split_sample_law(lite):
    """LITE_SAMPLER_IID against TTessel::split_sample (lib/ttessel.cpp:1783-1795, :1968-1991):
    n independent uniform positions along the cumulative halfedge length, sorted.
    (i) bit-exact against the numpy restatement of the draw + sort; (ii) the halfedge histogram
    against the exact multinomial law and against the oracle's own split_sample (two-sample
    chi-square); (iii) the spread of the split sum S0 = sum exp(theta.p) over repeated batches
    equals the spread over batches drawn by the oracle's split_sample, and the default stratified
    sampler has the same mean and a smaller spread (stated in DESIGN.md)."""
    import os
    import lite_oracle as lo
    import oracle_cpu as oc
    import philox_ref
    from util import GOLD
    soa_g, gold = load_golden("t100")
    window, segs = lo.parse_text(open(os.path.join(GOLD, "t100.txt")).read())
    t = lo.ttessel_from_segments(window, segs, lo.Rng(99))
    soa = lo.export_soa(t)
    cum, cum_he = inside_cum(soa)
    L = float(cum[-1])
    ctx = _ctx(lite, soa, ACS)
    ctx.sampler(True)
    n, seed, off = 20000, 0xC0FFEE, 4321
    ctx.add_splits_philox(n, seed, off)
    # (i) the draw and the sort, restated
    j = np.arange(n, dtype=np.uint64) + np.uint64(off)
    z = np.zeros(n, dtype=np.uint32)
    lo32, hi32 = (j & philox_ref.MASK).astype(np.uint32), (j >> np.uint64(32)).astype(np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    r1 = philox_ref.philox4x32_10(lo32, hi32, z + np.uint32(1), z, k0, k1)       # stream 1: the position
    sel = philox_ref.u53(r1[0], r1[1]) * L
    order = np.argsort(sel, kind="stable")
    r0 = philox_ref.philox4x32_10(lo32, hi32, z, z, k0, k1)                        # stream 0: x and U
    xs, us = philox_ref.u53(r0[0], r0[1]), philox_ref.u53(r0[2], r0[3])
    idx = np.minimum(np.searchsorted(cum, sel[order], side="right"), len(cum) - 1)
    he = ctx.fetch("SPLIT_E1")
    np.testing.assert_array_equal(he, cum_he[idx])
    np.testing.assert_array_equal(ctx.fetch("SPLIT_X"), xs[order])
    np.testing.assert_array_equal(ctx.fetch("SPLIT_U"), us[order])
    # (ii) histogram: multinomial(n, length / L) ...
    lens = soa["he_length"][cum_he]
    pos = {int(h): k for k, h in enumerate(cum_he)}
    counts = np.bincount([pos[int(h)] for h in he], minlength=len(cum_he))
    expect = n * lens / L
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    df = len(cum_he) - 1
    assert abs(chi2 - df) < 5 * math.sqrt(2 * df), (chi2, df)
    # ... and the oracle's own split_sample on the same tessellation
    hid = {id(h): i for i, h in enumerate(t.halfedges)}
    n_o = 6000
    o_he = np.array([hid[id(e)] for e in t.alt_length_weighted_random_halfedges(n_o)])
    o_counts = np.bincount([pos[int(h)] for h in o_he], minlength=len(cum_he))
    # pooled two-sample chi-square over halfedges with enough expected mass
    keep = expect * n_o / n > 3
    a, b = counts[keep].astype(float), o_counts[keep].astype(float)
    k1_, k2_ = math.sqrt(b.sum() / a.sum()), math.sqrt(a.sum() / b.sum())
    chi2_2 = float((((k1_ * a - k2_ * b) ** 2) / (a + b)).sum())
    df2 = int(keep.sum()) - 1
    assert abs(chi2_2 - df2) < 5 * math.sqrt(2 * df2), (chi2_2, df2)
    ctx.close()
    # (iii) spread of the split sum over repeated batches
    th = np.asarray(gold["pl_acs_1_theta"], dtype=np.float64)
    T = oc.Tess(soa)
    m, reps = 1500, 60

    def gpu_s0(iid):
        c2 = _ctx(lite, soa, ACS, record=False)
        c2.sampler(iid)
        c2.merges_flips()
        out = []
        for k in range(reps):
            c2.clear_splits()
            c2.add_splits_philox(m, 1234, k * m)
            lpl, _, _ = c2.pl_eval(th)
            out.append(lpl)
        c2.close()
        return np.array(out)
    lpl_iid, lpl_str = gpu_s0(True), gpu_s0(False)
    t.rnd = lo.Rng(2024)
    ms, fs = oc.merge_stats(T, ACS)["merge_stat"], oc.flip_stats(T, ACS)["flip_stat"]
    lpl_orc = []
    for k in range(reps):
        o_he, o_x, o_a = lo.draw_split_candidates(t, m)          # split_sample's draws, exact oracle
        ss = oc.split_stats(T, ACS, o_he, o_x, o_a)["split_stat"]
        lpl_orc.append(oc.pl(th, ss, fs, ms.sum(0), fs.sum(0), float(soa["int_length"]))[0])
    lpl_orc = np.array(lpl_orc)
    v_iid, v_str, v_orc = lpl_iid.var(ddof=1), lpl_str.var(ddof=1), lpl_orc.var(ddof=1)
    # same law: equal means (within 5 standard errors) and equal variances (F ratio, 60 vs 60 batches)
    se = math.sqrt(v_iid / reps + v_orc / reps)
    assert abs(lpl_iid.mean() - lpl_orc.mean()) < 5 * se
    assert 0.4 < v_iid / v_orc < 2.5, (v_iid, v_orc)
    # the stratified sampler: same mean, smaller spread
    assert abs(lpl_str.mean() - lpl_orc.mean()) < 5 * math.sqrt(v_str / reps + v_orc / reps)
    assert v_str < v_iid, (v_str, v_iid)


@pytest.mark.parametrize("what", ["twin", "list", "segment", "blocking", "range"])
def test_inconsistent_tessellation_is_caught_on_the_device(lite, what):
    """The export assertions run on the device behind the copy (k_validate): lite_tess_upload
    returns, no kernel dereferences an unchecked index, and the next synchronising call reports
    the inconsistency.  With LITE_OPT_VALIDATE_HOST the upload itself reports it.  The context
    stays usable."""
    soa, gold = load_golden("t100")
    bad = {k: np.array(v, copy=True) for k, v in soa.items()}
    if what == "twin":
        bad["he_twin"][5] = bad["he_twin"][9]
    elif what == "list":
        f = int(np.nonzero(bad["face_inside"])[0][3])
        o = int(bad["face_he_offset"][f])
        bad["face_he_list"][o], bad["face_he_list"][o + 1] = bad["face_he_list"][o + 1], bad["face_he_list"][o]
    elif what == "segment":
        bad["seg_he_start"][2] = len(bad["he_src"]) + 1000
    elif what == "blocking":
        bad["blocking"][0] = bad["non_blocking"][0]
    else:
        bad["he_src"][11] = 10 ** 9
    th = gold["pl_acs_1_theta"]
    ctx = lite.Context(0)
    ctx.energy_set(ACS)
    ctx.tess_upload(lite.TessSoA(bad))          # returns: the checks are queued behind the copy
    ctx.merges_flips()
    ctx.add_splits_philox(5000, 1, 0)
    with pytest.raises(lite.LiteError, match="inconsistent"):
        ctx.pl_eval(th)
    with pytest.raises(lite.LiteError, match="upload a tessellation"):
        ctx.merges_flips()
    ctx.validate_on_host(True)
    with pytest.raises(lite.LiteError):
        ctx.tess_upload(lite.TessSoA(bad))
    ctx.tess_upload(lite.TessSoA(soa))
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    assert ctx.pl_eval(th)[0] == pytest.approx(float(gold["pl_acs_1_lpl"]), rel=RTOL)
    ctx.close()


def test_random_sequence_of_good_and_bad_uploads(lite):
    """Sixty uploads in random order on one context -- consistent rectangles of three sizes, the
    window with holes, tessellations that fail on the host (offsets that do not tile, wrong
    int_length) and ones that fail on the device (broken twin, segment out of range) -- with the
    host-side validation switched on and off at random and nothing synchronised in between except
    by the calls themselves.  Every good upload evaluates to exactly what a fresh context gives;
    every bad one is reported, by the upload or by the next synchronising call, and leaves no
    tessellation behind."""
    import os
    rng = np.random.default_rng(20260)
    t100, _ = load_golden("t100")
    t11, _ = load_golden("t11")
    holed, _ = load_golden("holed")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_soa.npz"))
    c2 = {k: z[k] for k in z.files}
    c2["int_length"] = float(c2["int_length"])
    good = {"t100": (t100, 40000), "t11": (t11, 20000), "holed": (holed, 10000), "c2": (c2, 500000)}
    th = np.array([0.3, -0.02, 0.1, -0.5])

    def evaluate(c, n, seed):
        c.energy_set(ACS)
        c.merges_flips()
        c.add_splits_philox(n, seed, 0)
        return c.pl_eval(th)

    want = {}
    for name, (a, n) in good.items():
        c = lite.Context(0)
        c.tess_upload(lite.TessSoA(a))
        want[name] = evaluate(c, n, 7)
        c.close()

    def broken(kind):
        base = c2 if rng.random() < 0.5 else t100
        b = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in base.items()}
        if kind == "tiling":
            b["face_he_offset"][3] += 1
        elif kind == "length":
            b["int_length"] = float(b["int_length"]) * 1.5
        elif kind == "twin":
            b["he_twin"][5] = b["he_twin"][9]
        else:
            b["seg_he_start"][2] = len(b["he_src"]) + 1000
        return b

    ctx = lite.Context(0)
    n_good = n_bad = 0
    for it in range(60):
        on_host = bool(rng.random() < 0.3)
        ctx.validate_on_host(on_host)
        if rng.random() < 0.6:
            name = ["t100", "t11", "holed", "c2"][int(rng.integers(4))]
            a, n = good[name]
            ctx.tess_upload(lite.TessSoA(a))
            if rng.random() < 0.5:          # leave kernels queued, upload the same thing again
                ctx.energy_set(ACS)
                ctx.merges_flips()
                ctx.add_splits_philox(n, 99, 0)
                ctx.clear_splits()
                ctx.tess_upload(lite.TessSoA(a))
            v, g, H = evaluate(ctx, n, 7)
            wv, wg, wH = want[name]
            assert v == wv, (it, name)
            np.testing.assert_array_equal(g, wg)
            np.testing.assert_array_equal(H, wH)
            n_good += 1
        else:
            kind = ["tiling", "length", "twin", "segment"][int(rng.integers(4))]
            b = broken(kind)
            if kind in ("tiling", "length") or on_host:
                with pytest.raises(lite.LiteError):
                    ctx.tess_upload(lite.TessSoA(b))
            else:
                ctx.tess_upload(lite.TessSoA(b))        # the device finds it
                ctx.energy_set(ACS)
                ctx.merges_flips()
                ctx.add_splits_philox(1000, 1, 0)
                with pytest.raises(lite.LiteError, match="inconsistent"):
                    ctx.pl_eval(th)
            with pytest.raises(lite.LiteError, match="upload a tessellation"):
                ctx.merges_flips()
            n_bad += 1
    assert n_good > 20 and n_bad > 10
    ctx.close()


def test_holed_window_from_the_text_format_through_the_host_reader(lite):
    """The whole route a user of the reference takes with a tessellation of a window with holes:
    the text file (LineTes::write, lib/ttessel.cpp:1348-1370) -> lite_host_create_from_text ->
    lite_host_export -> lite_tess_upload -> pseudo-likelihood on the device, against the exact
    oracle on the tessellation that wrote the file (halfedges matched by their end points, the
    numbering of the two exports differs)."""
    import lite_oracle as lo
    from test_gen_face_cpu import grow
    from test_oracle_golden import holed_polygons
    t = grow(holed_polygons(), 13, 5, 90)
    h = lite.HostTess(text=lo.write_text(t))
    soa = h.export()
    a = soa.a
    ctx = lite.Context(0)
    ctx.tess_upload(soa)
    ctx.energy_set(ALL12)
    ctx.record_splits(True)
    nm, nf = ctx.merges_flips()
    assert nm == len(t.non_blocking_segments) and nf == 2 * len(t.blocking_segments)
    ctx.add_splits_philox(300, 19, 0)
    he, x, ang = ctx.fetch("SPLIT_E1"), ctx.fetch("SPLIT_X"), ctx.fetch("SPLIT_ANGLE")
    # the oracle's halfedge with the same end points
    okey = {}
    for k, hh in enumerate(t.halfedges):
        okey[(float(hh.src.p[0]), float(hh.src.p[1]), float(hh.tgt.p[0]), float(hh.tgt.p[1]))] = k

    def to_oracle(ids):
        src, dst = a["he_src"][ids], a["he_src"][a["he_twin"][ids]]
        return np.array([okey[(float(a["vx"][p]), float(a["vy"][p]), float(a["vx"][q]), float(a["vy"][q]))]
                         for p, q in zip(src, dst)], dtype=np.int32)
    res = lo.eval_candidates(t, lo.make_energy(lo.ALL12, None, t), to_oracle(he), x, ang)
    np.testing.assert_array_equal(to_oracle(ctx.fetch("SPLIT_E2")), res["split_e2"])
    scale = max(lo.face_area_2(lo.face2poly(f), t) for f in t.internal_faces())
    k_a2 = ALL12.index("face_area_2")
    keep = [k for k in range(len(ALL12)) if k != k_a2]
    got = ctx.fetch("SPLIT_STAT")
    _assert_stats("split", got[:, keep], res["split_stat"][:, keep], [ALL12[k] for k in keep])
    assert np.all(np.abs(got[:, k_a2] - res["split_stat"][:, k_a2]) <= ATOL_STAT + RTOL * np.abs(res["split_stat"][:, k_a2]) + 16 * 2.3e-16 * scale)
    # merges and flips come in the order of the two lists, which the file does not fix: compare as sets of rows
    for name, want in (("MERGE_STAT", res["merge_stat"]), ("FLIP_STAT", res["flip_stat"])):
        g = ctx.fetch(name)
        assert g.shape == want.shape
        gs, ws = g[np.lexsort(np.round(g, 6).T[::-1])], want[np.lexsort(np.round(want, 6).T[::-1])]
        assert np.all(np.abs(gs[:, keep] - ws[:, keep]) <= 1e-9 + RTOL * np.abs(ws[:, keep]))
        assert np.all(np.abs(gs[:, k_a2] - ws[:, k_a2]) <= 1e-9 + RTOL * np.abs(ws[:, k_a2]) + 16 * 2.3e-16 * scale)
    th = 0.01 * np.random.default_rng(3).standard_normal(len(ALL12))
    th[k_a2] *= 1e-4
    lpl, g, H = ctx.pl_eval(th)
    wl, wg, wH = lo.pl_from_stats(th, res["split_stat"], res["merge_stat"], res["flip_stat"], float(soa.int_length))
    assert lpl == pytest.approx(wl, rel=RTOL)
    np.testing.assert_allclose(g, wg, rtol=RTOL, atol=1e-9 * max(1.0, np.abs(wg).max()))
    np.testing.assert_allclose(H, wH, rtol=RTOL, atol=1e-9 * max(1.0, np.abs(wH).max()))
    ctx.close()


def test_golden_window_with_holes(lite):
    """Committed golden vectors for the reference's holed domain (tests/golden/holed.npz, made by
    the exact oracle, make_golden_holed.py): fed split list, every merge and flip, all twelve
    features, PL value / gradient / Hessian — no oracle at run time."""
    soa, gold = load_golden("holed")
    ctx = _ctx(lite, soa, ALL12)
    ctx.merges_flips()
    ctx.add_splits_given(gold["cand_he"], gold["cand_x"], gold["cand_angle"])
    np.testing.assert_array_equal(ctx.fetch("SPLIT_E2"), gold["split_e2"])
    np.testing.assert_array_equal(ctx.fetch("MERGE_HE"), gold["merge_he"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_E2"), gold["flip_e2"])
    np.testing.assert_array_equal(ctx.fetch("FLIP_RIGHT"), gold["flip_right"])
    k_a2, scale = ALL12.index("face_area_2"), float(gold["area2_scale"])
    keep = [k for k in range(len(ALL12)) if k != k_a2]
    for name, key in (("SPLIT_STAT", "split_stat"), ("MERGE_STAT", "merge_stat"), ("FLIP_STAT", "flip_stat")):
        got, want = ctx.fetch(name), gold[key]
        _assert_stats(name, got[:, keep], want[:, keep], [ALL12[k] for k in keep])
        assert np.all(np.abs(got[:, k_a2] - want[:, k_a2]) <= ATOL_STAT + RTOL * np.abs(want[:, k_a2]) + 16 * 2.3e-16 * scale)
    lpl, g, H = ctx.pl_eval(gold["pl_theta"])
    assert lpl == pytest.approx(float(gold["pl_lpl"]), rel=RTOL)
    np.testing.assert_allclose(g, gold["pl_grad"], rtol=RTOL, atol=1e-9 * max(1.0, np.abs(gold["pl_grad"]).max()))
    np.testing.assert_allclose(H, gold["pl_hess"], rtol=RTOL, atol=1e-9 * max(1.0, np.abs(gold["pl_hess"]).max()))
    assert ctx.pl_failures() == (0, 0)
    ctx.close()
