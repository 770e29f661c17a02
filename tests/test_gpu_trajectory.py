"""Trajectory parity (north_star): seeded runs of the C++ host mirror (System::mc / SimulationControl::PI_nvt_mc over the GPU
engine) must reproduce the reference's accept/reject trajectory — move type and decision of every step for the first 10^4 moves —
and the trial energies along the way to 1e-10.  Golden trajectories come from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from mpmcxx_b200 import workloads as W
from tests import cases

pytestmark = pytest.mark.gpu


def _subterm_scale(s):
    """Sum of |energy sub-terms| of the start configuration: what the total energy of a step is a (cancelling) sum of."""
    from mpmcxx_b200 import engine
    t = s.copy()
    t.opts["ensemble"] = "nvt"
    e = engine.Engine(t)
    o = e.energy()
    e.close()
    return sum(abs(o[k]) for k in ("rd_pair", "rd_lrc_pair", "rd_lrc_self", "es_real", "es_self_intra", "es_reciprocal", "es_self", "polarization_energy"))


def _n_mobile(s):
    return float(len(set(int(m) for m, f in zip(s.mol, s.frozen) if not f)))


@pytest.mark.parametrize("name", sorted(cases.TRAJ))
def test_trajectory_matches_reference(name, tmp_path):
    from mpmcxx_b200 import host_binding
    s, r = cases.load_golden_traj(name)
    P, ref = int(r["P"]), r["traj"]
    inp = W.write_reference_job(s, str(tmp_path))
    log, summary = host_binding.run(inp, P=P, max_steps=len(ref), capacity=len(ref))
    assert len(log) == len(ref)
    same_move = log[:, 0] == ref[:, 0]
    same_acc = log[:, 3] == ref[:, 3]
    first_bad = int(np.argmin(same_move & same_acc)) if not (same_move & same_acc).all() else -1
    assert first_bad == -1, "trajectory diverges at step %d: ours %s reference %s" % (first_bad, log[first_bad], ref[first_bad])
    fin = np.isfinite(ref[:, 1]) & (np.abs(ref[:, 1]) < 1e30)
    # 1e-10 of the scale of what is summed: the total cancels ~1e5 K of Ewald sub-terms into a few K, so the scale is the
    # sum of |sub-terms| of the start configuration (evaluated through the engine), per molecule where N changes along the chain
    scale = _subterm_scale(s) * (np.maximum(ref[fin, 4], 1.0) / max(_n_mobile(s), 1.0) if not P else 1.0)
    assert (np.abs(log[fin, 1] - ref[fin, 1]) <= 1e-10 * np.maximum(scale, np.abs(ref[fin, 1]))).all()
    bfin = np.isfinite(ref[:, 2])          # (with an orientational term the reference's factor is exp(+-1e26): 0 or +inf, matched as such)
    assert (log[~bfin, 2] == ref[~bfin, 2]).all()
    bf_scale = np.maximum(np.abs(ref[bfin, 2]), 1e-300)
    ok = np.abs(log[bfin, 2] - ref[bfin, 2]) / bf_scale < 1e-6
    assert ok.all()
    assert summary[6] == ref[:, 3].sum() and summary[7] == len(ref) - ref[:, 3].sum()
    if P:   # path integrals: the kinetic estimator after every step as well
        assert np.allclose(log[:, 4], ref[:, 4], rtol=1e-10, atol=0)


def test_host_reader_and_energy_match_reference_fixture(tmp_path):
    """PQR + input-file reader -> flatten -> engine equals the reference's energy() on the 104-site known answer."""
    from mpmcxx_b200 import host_binding
    s, r = cases.load_golden("kat_gs_ranked_palmo")
    s.opts.update({"seed": "1", "numsteps": "1"})
    inp = W.write_reference_job(s, str(tmp_path))
    o = host_binding.energy(inp)
    e = r["ref_energy_cold"]
    assert abs(o["rd"] - e[1]) < 1e-10 * abs(e[1]) and abs(o["polar"] - e[3]) < 1e-10 * abs(e[3])
    assert o["iterations"] == int(r["ref_iterations"])


@pytest.mark.parametrize("name", sorted(cases.SHIPPED))
def test_shipped_sample_directory_runs_unmodified(name, tmp_path):
    """BASELINE config 1 (sample-input/pi000-free-argon-2K, `-P 8 equilibrate.in`) with the files as shipped, through the mirror's
    readers and the GPU engine: accept/reject decisions and the kinetic-energy series of the reference for 10^4 steps."""
    import os
    from mpmcxx_b200 import host_binding
    files, inp, P, ref = cases.load_shipped(name)
    cases.write_shipped(files, str(tmp_path))
    log, summary = host_binding.run(os.path.join(str(tmp_path), inp), P=P, max_steps=len(ref), capacity=len(ref))
    assert len(log) == len(ref)
    same = (log[:, 0] == ref[:, 0]) & (log[:, 3] == ref[:, 3])
    assert same.all(), "trajectory diverges at step %d" % int(np.argmin(same))
    assert np.allclose(log[:, 4], ref[:, 4], rtol=1e-10, atol=0)
    assert (np.abs(log[:, 1] - ref[:, 1]) <= 1e-10 * np.maximum(np.abs(ref[:, 1]), 1.0)).all()
    assert summary[6] == ref[:, 3].sum()
