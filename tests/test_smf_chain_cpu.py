"""The SMF chain source the GPU kernel runs (lite_b200/csrc/smf_chain.h), compiled for
the host by tests/cpp/smf_host.cpp and replayed proposal by proposal in the exact
oracle: crossed edges and end points, Hastings ratios (lib/ttessel.cpp:3158-3252),
energy variations and the blocking / non-blocking bookkeeping of update() must
agree.  The harness also runs the consistency check after every step."""
import math
import os
import subprocess

import numpy as np
import pytest

from smf_replay import TRACE_DT, replay
from lite_b200.capi import FEATURE_IDS

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "smf_host")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="tests/cpp/smf_host not built (make)")

ENERGIES = {
    # applications/checkACS.cpp:113-117 with tau = 6
    "checkacs": (["is_segment_internal", "edge_length"], [-math.log(6.0), 5.0 / math.pi]),
    # demos/sim_acs.cpp:64-76 (tau=12 there; 4 here so that merges and flips get rejected too)
    "acs": (["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"],
            [3.0 / math.pi, 4.0 ** 4 * 0.05, 1.0, -math.log(4.0)]),
    # every feature at once, the two face features that depend on the polygon start included
    "all": (["is_point_inside_window", "edge_length", "edge_length_internal", "seg_number",
             "is_segment_internal", "segment_size_2", "face_number", "face_area_2", "face_perimeter",
             "face_shape", "face_sum_of_angles", "min_angle"],
            [0.1, 0.7, 0.3, -0.4, -1.2, 0.01, -0.2, 3.0, 0.05, 0.02, 0.5, 0.4]),
}


def run_host(tmp_path, cap, n_steps, seed, chain, names, thetas, ps=0.33, pm=0.33):
    out = str(tmp_path / "trace.bin")
    cmd = [EXE, str(cap), str(n_steps), str(seed), str(chain), repr(ps), repr(pm), out, str(len(names))]
    for n, th in zip(names, thetas):
        cmd += [str(FEATURE_IDS[n]), repr(float(th))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    words = res.stdout.split()
    return np.fromfile(out, dtype=TRACE_DT), int(words[1]), int(words[3]), [float(w) for w in words[4:]]


@pytest.mark.parametrize("tag,n_steps,seed", [("checkacs", 260, 5), ("acs", 260, 7), ("all", 200, 11)])
def test_chain_replays_in_exact_oracle(tmp_path, tag, n_steps, seed):
    names, thetas = ENERGIES[tag]
    trace, done, status, row = run_host(tmp_path, 256, n_steps, seed, 3, names, thetas)
    assert done == n_steps and status == 0
    t, evaluated = replay(trace, names, thetas, 0.33, 0.33)
    assert all(n > 10 for n in evaluated), evaluated
    # the checkACS columns (applications/checkACS.cpp:136-148) of the final state
    assert row[1] == len(t.all_segments) - 4
    assert row[2] == len(t.blocking_segments) and row[3] == len(t.non_blocking_segments)
    assert row[0] == pytest.approx((t.int_length - 4.0) / 2, rel=1e-11)
    n_boundary_vertices = sum(s.number_of_edges() for s in t.all_segments[:4])
    assert row[4] == len(t.vertices) - n_boundary_vertices


def test_capacity_is_reported_not_ignored(tmp_path):
    names, thetas = ENERGIES["checkacs"]
    trace, done, status, row = run_host(tmp_path, 24, 4000, 5, 0, names, thetas)
    assert status == 1 and done < 4000      # smf::ST_CAPACITY, the chain stops
    assert row[1] <= 24 - 4


def test_long_chain_stays_consistent_and_near_equilibrium(tmp_path):
    """30 000 proposals of the checkACS model (tau = 6), consistency check after each:
    detailed balance between splits and merges gives E[n_nb] = tau * E[int_length] / pi."""
    names, thetas = ENERGIES["checkacs"]
    trace, done, status, row = run_host(tmp_path, 1024, 30000, 5, 1, names, thetas)
    assert done == 30000 and status == 0
    tail = trace[15000:]
    est = tail["n_nb"].mean() * math.pi / tail["int_length"].mean()
    assert abs(est - 6.0) < 0.6, est


def test_chain_source_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """compute-sanitizer is not available on the GPU pool; the chain logic is one source for
    host and device, so its memory safety is checked on the host build: a long chain, all
    twelve features, and a chain that runs into its capacity limit."""
    exe = str(tmp_path / "smf_host_asan")
    src = os.path.join(HERE, "cpp", "smf_host.cpp")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                         "-o", exe, src], capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here: " + cc.stderr[-200:])
    trace = str(tmp_path / "t.bin")
    names, thetas = ENERGIES["all"]
    feats = []
    for n, th in zip(names, thetas):
        feats += [str(FEATURE_IDS[n]), repr(float(th))]
    runs = [
        ["1024", "12000", "5", "1", "0.33", "0.33", trace, "2", "3", repr(-math.log(6.0)), "1", repr(5 / math.pi)],
        ["300", "6000", "7", "2", "0.33", "0.33", trace, str(len(names))] + feats,
        ["24", "4000", "5", "0", "0.33", "0.33", trace, "1", "3", repr(-math.log(6.0))],
    ]
    for args in runs:
        res = subprocess.run([exe] + args, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        assert "runtime error" not in res.stderr and "AddressSanitizer" not in res.stderr, res.stderr[-2000:]
