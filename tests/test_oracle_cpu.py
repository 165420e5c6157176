"""The C++ double restatement (oracle/oracle_cpu.cpp: the timed CPU baseline)
pinned against the exact oracle's committed golden vectors, so that the number
bench.py reports next to the GPU is the time of a *correct* CPU path."""
import numpy as np
import pytest

import oracle_cpu as oc
from util import ATOL_STAT, RTOL, inside_cum, load_golden

ALL12 = ["is_point_inside_window", "edge_length", "face_number", "face_area_2",
         "face_perimeter", "face_shape", "face_sum_of_angles", "min_angle",
         "seg_number", "is_segment_internal", "minus_is_segment_internal", "segment_size_2"]
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]


def _close(name, got, want):
    tol = ATOL_STAT + RTOL * np.abs(want)
    bad = np.argwhere(np.abs(got - want) > tol)
    assert len(bad) == 0, "%s: %d entries differ, first %s got %r want %r" % (
        name, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("name", ["t11", "t100"])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_port_matches_exact_oracle(name, nthreads):
    soa, gold = load_golden(name)
    t = oc.Tess(soa)
    r = oc.split_stats(t, ALL12, gold["cand_he"], gold["cand_x"], gold["cand_angle"], nthreads)
    assert r["bad"] == 0
    np.testing.assert_array_equal(r["split_e2"], gold["split_e2"])
    np.testing.assert_allclose(r["split_p1"], gold["split_p1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(r["split_p2"], gold["split_p2"], rtol=0, atol=1e-12)
    _close("split", r["split_stat"], gold["split_stat"])
    m = oc.merge_stats(t, ALL12, nthreads)
    np.testing.assert_array_equal(m["merge_he"], gold["merge_he"])
    _close("merge", m["merge_stat"], gold["merge_stat"])
    f = oc.flip_stats(t, ALL12, nthreads)
    assert f["bad"] == 0
    np.testing.assert_array_equal(f["flip_e1"], gold["flip_e1"])
    np.testing.assert_array_equal(f["flip_e2"], gold["flip_e2"])
    np.testing.assert_array_equal(f["flip_right"], gold["flip_right"])
    np.testing.assert_allclose(f["flip_p2"], gold["flip_p2"], rtol=0, atol=1e-12)
    _close("flip", f["flip_stat"], gold["flip_stat"])


@pytest.mark.parametrize("name", ["t11", "t100"])
def test_port_pl_value_gradient_hessian(name):
    soa, gold = load_golden(name)
    cols = [ALL12.index(f) for f in ACS]
    ss, ms, fs = (gold[k][:, cols] for k in ("split_stat", "merge_stat", "flip_stat"))
    for k in (0, 1):
        th = gold["pl_acs_%d_theta" % k]
        for nt in (1, 4):
            v, g, H = oc.pl(th, ss, fs, ms.sum(0), fs.sum(0), float(soa["int_length"]), nthreads=nt)
            wl, wg, wH = (gold["pl_acs_%d_%s" % (k, s)] for s in ("lpl", "grad", "hess"))
            assert v == pytest.approx(float(wl), rel=RTOL)
            np.testing.assert_allclose(g, wg, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wg).max()))
            np.testing.assert_allclose(H, wH, rtol=RTOL, atol=1e-12 * max(1.0, np.abs(wH).max()))


def test_port_sampler_is_the_device_sampler():
    """oracle_sample_splits == the numpy restatement of the device's Philox sampler."""
    import philox_ref
    soa, _ = load_golden("t100")
    t = oc.Tess(soa)
    n, seed, off = 5000, 0x5EED, 77
    he, x, ang = oc.sample_splits(t, n, seed, off, nthreads=2)
    cum, cum_he = inside_cum(soa)
    he2, x2, U2 = philox_ref.sample_splits(cum, cum_he, n, seed, off, 0, n)
    np.testing.assert_array_equal(he, he2)
    np.testing.assert_array_equal(x, x2)
    np.testing.assert_allclose(ang, np.arccos(1 - 2 * U2), rtol=0, atol=1e-15)
    # a shard of the batch is the same slice
    he3, x3, _ = oc.sample_splits(t, n, seed, off, j0=1000, n=500)
    np.testing.assert_array_equal(he3, he[1000:1500])
    np.testing.assert_array_equal(x3, x[1000:1500])


def test_port_whole_pass_crtt_closed_form():
    """lpl = th*n_nb - (u/pi) e^th - 2 n_b for the CRTT energy whatever the sample
    (reference demos/pseudo_lik_inference.cpp:80-81)."""
    import math
    soa, _ = load_golden("t100")
    t = oc.Tess(soa)
    u = float(soa["int_length"])
    for th in (-0.3, 1.1):
        v, g, H = oc.pl_pass(t, ["minus_is_segment_internal"], [th], 3000, 11, nthreads=2)
        assert v == pytest.approx(th * t.n_nb - (u / math.pi) * math.exp(th) - 2 * t.n_b, rel=1e-12)
        assert g[0] == pytest.approx(t.n_nb - (u / math.pi) * math.exp(th), rel=1e-12)
        assert H[0, 0] == pytest.approx(-(u / math.pi) * math.exp(th), rel=1e-12)


def test_port_lines_clip_checksum_is_the_checksum_of_its_parts():
    """oracle_lines_clip (config C4: sample + clip, nothing stored) against
    sample_splits + the Split constructor run separately."""
    soa, _ = load_golden("t100")
    t = oc.Tess(soa)
    n, seed, off = 50000, 99, 5
    cs, bad = oc.lines_clip(t, n, seed, off, nthreads=3)
    he, x, ang = oc.sample_splits(t, n, seed, off, nthreads=2)
    r = oc.split_stats(t, ["face_number"], he, x, ang, nthreads=2)
    assert bad == r["bad"] == 0
    key = (he.astype(np.uint64) << np.uint64(32)) | r["split_e2"].astype(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        key = key * np.uint64(0x9E3779B97F4A7C15)
        key ^= key >> np.uint64(29)
    assert int(np.bitwise_xor.reduce(key)) == int(cs[0])
    assert int(r["split_e2"].astype(np.uint32).astype(np.uint64).sum()) == int(cs[1])
    # a shard of the batch: counters are global, so two halves combine to the whole
    a, _ = oc.lines_clip(t, n, seed, off, 0, n // 2, 2)
    b, _ = oc.lines_clip(t, n, seed, off, n // 2, n - n // 2, 2)
    assert int(a[0]) ^ int(b[0]) == int(cs[0]) and (int(a[1]) + int(b[1])) % 2 ** 64 == int(cs[1])
