"""GPU tests of the SMF chain ensemble (BASELINE.json config 5)
through the C ABI (lite_smf_*).

Parity: a device chain's trace is replayed proposal by proposal in the exact oracle
(tests/smf_replay.py): crossed-edge identity and end points, Hastings ratios and
energy variations to 1e-9 relative, segment bookkeeping exact.  At ensemble scale the
checks are the size-independent properties the chain offers: proposal counts, the
consistency check of every chain, independence of the sharding, and the detailed
balance identity E[n_nb] = tau E[int_length] / pi of the stationary law."""
import math
import os
import subprocess

import numpy as np
import pytest

import lite_b200
from lite_b200.capi import FEATURE_IDS, SMF_COLS
from smf_replay import TRACE_DT, replay
from test_smf_chain_cpu import ENERGIES, EXE

pytestmark = pytest.mark.gpu

COL = {n: i for i, n in enumerate(SMF_COLS)}


@pytest.fixture(scope="module")
def ctx():
    c = lite_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("tag,n_steps,seed", [("checkacs", 260, 5), ("acs", 260, 7), ("all", 200, 11)])
def test_device_chain_replays_in_exact_oracle(ctx, tag, n_steps, seed):
    names, thetas = ENERGIES[tag]
    ctx.smf_create(64, 252, seed=seed)
    ctx.smf_set_model(names, thetas, 0.33, 0.33)
    raw = ctx.smf_trace(3, n_steps)
    assert len(raw) == n_steps
    trace = np.frombuffer(raw.tobytes(), dtype=TRACE_DT)
    t, evaluated = replay(trace, names, thetas, 0.33, 0.33)
    assert all(n > 10 for n in evaluated), evaluated
    rows = ctx.smf_stats(validate=True)
    assert rows[COL["n"], 3] == len(t.all_segments) - 4
    assert rows[COL["n_bl"], 3] == len(t.blocking_segments)
    assert rows[COL["n_nbl"], 3] == len(t.non_blocking_segments)
    assert rows[COL["l"], 3] == pytest.approx((t.int_length - 4.0) / 2, rel=1e-11)
    assert rows[COL["m"], 3] == len(t.vertices) - sum(s.number_of_edges() for s in t.all_segments[:4])
    assert np.all(rows[COL["n"], :3] == 0)  # the other chains did not move
    ctx.smf_destroy()


@pytest.mark.skipif(not os.path.exists(EXE), reason="tests/cpp/smf_host not built")
def test_device_chain_equals_host_build_of_the_same_source(ctx, tmp_path):
    """Same seed, same global chain id: the kernel and the host build of smf_chain.h
    make the same proposals and the same decisions (floating point to 1e-9)."""
    names, thetas = ENERGIES["acs"]
    n_steps, seed, chain = 3000, 21, 17
    out = str(tmp_path / "trace.bin")
    cmd = [EXE, "512", str(n_steps), str(seed), str(chain), "0.33", "0.33", out, str(len(names))]
    for n, th in zip(names, thetas):
        cmd += [str(FEATURE_IDS[n]), repr(float(th))]
    subprocess.run(cmd, check=True, capture_output=True)
    host = np.fromfile(out, dtype=TRACE_DT)
    ctx.smf_create(32, 508, seed=seed)
    ctx.smf_set_model(names, thetas, 0.33, 0.33)
    dev = np.frombuffer(ctx.smf_trace(chain, n_steps).tobytes(), dtype=TRACE_DT)
    assert len(dev) == len(host) == n_steps
    for f in ("type", "status", "n_seg", "n_nb", "n_b"):
        assert np.array_equal(dev[f], host[f]), f
    for f in ("r", "e_var", "x", "angle", "int_length"):
        assert np.allclose(dev[f], host[f], rtol=1e-9, atol=1e-12), f
    for f in ("p1", "p2", "e1s", "e1t", "e2s", "e2t"):
        assert np.allclose(dev[f], host[f], rtol=0, atol=1e-12), f
    ctx.smf_destroy()


def test_ensemble_rows_counts_and_shard_independence(ctx):
    names, thetas = ENERGIES["checkacs"]
    n, n_samples, nipl = 3000, 4, 250
    ctx.smf_create(n, 300, seed=5)
    ctx.smf_set_model(names, thetas, 0.33, 0.33)
    rows = ctx.smf_run(n_samples, nipl)
    assert rows.shape == (n_samples, len(SMF_COLS), n)
    prop = rows[:, COL["prop_S"]] + rows[:, COL["prop_M"]] + rows[:, COL["prop_F"]]
    assert np.all(prop == nipl)                      # ModifCounts of each step(nipl) call
    assert np.all(rows[:, COL["acc_S"]] <= rows[:, COL["prop_S"]])
    assert np.all(rows[:, COL["n"]] == rows[:, COL["n_bl"]] + rows[:, COL["n_nbl"]])
    # n(T) moves by accepted splits minus accepted merges
    dn = np.diff(np.concatenate([np.zeros((1, n)), rows[:, COL["n"]]]), axis=0)
    assert np.array_equal(dn, rows[:, COL["acc_S"]] - rows[:, COL["acc_M"]])
    # checkACS energy as written: theta_seg * n(T) + (tau-1)/pi * (change of the boundary length ~ 0)
    assert np.allclose(rows[-1, COL["energy"]], thetas[0] * rows[-1, COL["n"]], atol=1e-9)
    assert np.array_equal(ctx.smf_stats(validate=True)[:5], rows[-1, :5])
    # chains are identified by their global id: a different ensemble size (hence
    # another sharding) does not change chain i
    ctx.smf_create(500, 300, seed=5)
    ctx.smf_set_model(names, thetas, 0.33, 0.33)
    rows2 = ctx.smf_run(n_samples, nipl)
    assert np.array_equal(rows2, rows[:, :, :500])
    # different seeds give different chains, and chains differ from one another
    assert len(np.unique(rows[-1, COL["l"]])) > 0.99 * n
    ctx.smf_destroy()


def test_capacity_overflow_is_an_error(ctx):
    names, thetas = ENERGIES["checkacs"]
    ctx.smf_create(256, 16, seed=5)
    ctx.smf_set_model(names, thetas, 0.33, 0.33)
    with pytest.raises(lite_b200.LiteError, match="max_segments"):
        ctx.smf_run(1, 4000, rows=False)
    assert ctx.smf_stats()[COL["n"]].max() <= 16
    ctx.smf_destroy()


def test_ensemble_detailed_balance_at_equilibrium(ctx):
    """tau = 3 (about 150 cells): after burn-in the split and merge fluxes balance,
    E[n_nb] = tau * E[int_length] / pi, whatever the dummy details of the sampler."""
    tau = 3.0
    names, thetas = ["is_segment_internal", "edge_length"], [-math.log(tau), (tau - 1) / math.pi]
    n = 4096
    ctx.smf_create(n, 400, seed=9)
    ctx.smf_set_model(names, thetas, 0.33, 0.33)
    ctx.smf_run(1, 6000, rows=False)
    rows = ctx.smf_run(4, 500)
    nnb = rows[:, COL["n_nbl"]].mean()
    int_length = (2 * rows[:, COL["l"]] + 4.0).mean()
    est = nnb * math.pi / int_length
    assert abs(est - tau) < 0.02 * tau, est
    # acceptance rates of the three move types are all strictly between 0 and 1
    for a, p in (("acc_S", "prop_S"), ("acc_M", "prop_M"), ("acc_F", "prop_F")):
        rate = rows[:, COL[a]].sum() / rows[:, COL[p]].sum()
        assert 0.5 < rate <= 1.0, (a, rate)
    ctx.smf_stats(validate=True)
    ctx.smf_destroy()


def test_exported_chain_feeds_the_pseudo_likelihood_path(ctx):
    """A chain's tessellation comes back as a host model, passes its checks, and goes
    through lite_tess_upload + the CRTT closed form (demos/pseudo_lik_inference.cpp:80-81)."""
    tau = 4.0
    ctx.smf_create(16, 600, seed=3)
    ctx.smf_set_model(["minus_is_segment_internal"], [math.log(tau)], 0.33, 0.33)
    rows = ctx.smf_run(1, 12000)
    h = ctx.smf_export(7)
    h.check()
    cells, nnb, nb = h.counts()
    assert cells == rows[0, COL["n"], 7] + 1 and nnb == rows[0, COL["n_nbl"], 7] and nb == rows[0, COL["n_bl"], 7]
    soa = h.export()
    assert soa.int_length == pytest.approx(2 * rows[0, COL["l"], 7] + 4.0, rel=1e-12)
    ctx.tess_upload(soa)
    ctx.energy_set(["minus_is_segment_internal"])
    ctx.merges_flips()
    ctx.clear_splits()
    ctx.add_splits_philox(20000, 1, 0)
    th = math.log(nnb * math.pi / soa.int_length)
    lpl, g, H = ctx.pl_eval([th])
    assert abs(g[0]) < 1e-9 * nnb                      # the closed-form estimate is the optimum
    assert H[0, 0] == pytest.approx(-nnb, rel=1e-9)
    assert abs(th - math.log(tau)) < 0.3
    # and the host model can carry on from where the device chain stopped
    h.crtt_run(tau, 2000, 1)
    h.check()
    ctx.smf_destroy()


def test_simulate_gibbs_model_then_recover_theta_by_pl_newton(ctx):
    """End to end over both device paths: tessellations are simulated on the GPU under a
    Gibbs energy with a face interaction (theta_angles * face_sum_of_angles + theta_seg *
    is_segment_internal, the angle and segment terms of demos/sim_acs.cpp:64-76), brought
    back, and the pseudo-likelihood path (SetEnergy + AddSplits + damped Newton on
    value / gradient / Hessian, lib/ttessel.cpp:3454-3505) must recover the parameters
    that generated them, on average over independent chains."""
    tau, th_ang = 5.0, 1.0
    names = ["face_sum_of_angles", "is_segment_internal"]
    truth = np.array([th_ang, -math.log(tau)])
    n_chains = 24
    ctx.smf_create(n_chains, 1000, seed=77)
    ctx.smf_set_model(names, truth, 0.33, 0.33)
    rows = ctx.smf_run(2, 30000)
    ctx.smf_stats(validate=True)
    n_end = rows[-1, COL["n"]]
    assert n_end.min() > 30                      # a real tessellation, not the empty window
    # stationarity: the second half of the run does not change the mean size much
    assert abs(rows[1, COL["n"]].mean() - rows[0, COL["n"]].mean()) < 0.15 * rows[0, COL["n"]].mean()
    est = []
    for k in range(n_chains):
        h = ctx.smf_export(k)
        soa = h.export()
        ctx.tess_upload(soa)
        ctx.energy_set(names)
        ctx.merges_flips()
        ctx.clear_splits()
        ctx.add_splits_philox(200000, 1000 + k, 0)
        th = np.zeros(2)
        lpl, g, H = ctx.pl_eval(th)
        for _ in range(50):                      # damped Newton on a concave function
            step = np.linalg.solve(H, g)
            lam = 1.0
            while True:
                cand = th - lam * step
                l2, g2, H2 = ctx.pl_eval(cand)
                if l2 >= lpl or lam < 1e-6:
                    break
                lam *= 0.5
            th, lpl, g, H = cand, l2, g2, H2
            if np.abs(g).max() < 1e-7 * max(1.0, abs(lpl)):
                break
        assert np.abs(g).max() < 1e-5 * max(1.0, abs(lpl)), "Newton did not converge on chain %d" % k
        assert np.all(np.linalg.eigvalsh(H) < 0)
        est.append(th)
    est = np.array(est)
    mean, sem = est.mean(0), est.std(0, ddof=1) / math.sqrt(n_chains)
    # within 4 standard errors of the truth (plus a small allowance for the bias of the
    # maximum pseudo-likelihood estimator on ~100-cell tessellations)
    assert np.all(np.abs(mean - truth) < 4 * sem + 0.05 * np.abs(truth)), (mean, sem, truth)
    ctx.smf_destroy()


def test_intersection_output_of_checkacs(ctx):
    """lite_smf_intersections = the inter file of applications/checkACS.cpp:151-172: every
    internal edge against the four window edges and the two mid-lines.  Checked against a plain
    numpy computation on the exported chains: a mid-line is crossed by the edges that straddle
    it, a window edge is touched by the internal edges that end on it."""
    ctx.smf_create(40, 300, seed=9)
    ctx.smf_set_model(["is_segment_internal", "edge_length"], [-math.log(4.0), 3.0 / math.pi])
    ctx.smf_run(1, 1500, rows=False)
    probes = np.array([[0, 0, 1, 0], [1, 0, 1, 1], [1, 1, 0, 1], [0, 1, 0, 0], [0, .5, 1, .5], [.5, 0, .5, 1]], dtype=np.float64)
    ch, kk, xy = ctx.smf_intersections(probes)
    assert len(ch) > 40 * 10
    for chain in (0, 7, 39):
        a = ctx.smf_export(chain).export().a
        inside = a["face_inside"][a["he_face"]].astype(bool)
        internal = inside & inside[a["he_twin"]] & (np.arange(len(inside)) < a["he_twin"])      # one halfedge per internal edge
        src, dst = a["he_src"][internal], a["he_src"][a["he_twin"][internal]]
        ax, ay, bx, by = a["vx"][src], a["vy"][src], a["vx"][dst], a["vy"][dst]
        want = []
        for k, (x0, y0, x1, y1) in enumerate(probes):
            if k < 4:       # window edge: touched by an end point lying on it
                for px, py in ((ax, ay), (bx, by)):
                    on = (px == x0) if x0 == x1 else (py == y0)
                    want += [(k, float(u), float(v)) for u, v in zip(px[on], py[on])]
            elif k == 4:    # y = 1/2
                cr = (ay - .5) * (by - .5) < 0
                t = (ay[cr] - .5) / (ay[cr] - by[cr])
                want += [(k, float(u), .5) for u in ax[cr] + t * (bx[cr] - ax[cr])]
            else:           # x = 1/2
                cr = (ax - .5) * (bx - .5) < 0
                t = (ax[cr] - .5) / (ax[cr] - bx[cr])
                want += [(k, .5, float(v)) for v in ay[cr] + t * (by[cr] - ay[cr])]
        got = [(int(k), float(x), float(y)) for k, x, y in zip(kk[ch == chain], xy[ch == chain, 0], xy[ch == chain, 1])]
        assert len(got) == len(want), (chain, len(got), len(want))
        for g, w in zip(sorted(got), sorted(want)):
            assert g[0] == w[0] and abs(g[1] - w[1]) < 1e-12 and abs(g[2] - w[2]) < 1e-12
