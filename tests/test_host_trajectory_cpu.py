"""Trajectory parity of the C++ host mirror WITHOUT a GPU: the mirror's Monte Carlo drivers (System::mc, SimulationControl::PI_nvt_mc:
move generation, RNG order, Boltzmann factors, accept/reject, restore, uVT insert/remove, bead moves) linked against a test-only
shim that answers the engine's C-ABI from the CPU oracle (tests/shim/oracle_engine.c), compared with the reference's golden
trajectories step by step.  The GPU tests (tests/test_gpu_trajectory.py) run the same drivers over the real engine."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mpmcxx_b200 import workloads as W
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mpmcxx_b200", "host")


@pytest.fixture(scope="module")
def host_cpu(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("host_cpu"))
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle")]
    cflags = ["-O2", "-std=c11", "-fPIC", "-fopenmp", "-ffp-contract=off"]
    subprocess.run(["gcc"] + cflags + inc + ["-c", os.path.join(ROOT, "oracle", "oracle.c"), "-o", os.path.join(d, "oracle.o")], check=True)
    subprocess.run(["gcc"] + cflags + inc + ["-c", os.path.join(ROOT, "tests", "shim", "oracle_engine.c"), "-o", os.path.join(d, "shim.o")], check=True)
    lib = os.path.join(d, "libmpmc_host_cpu.so")
    srcs = [os.path.join(HOST, f) for f in ("host.cpp", "sim_control.cpp", "host_capi.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp"] + inc + ["-o", lib] + srcs +
                   [os.path.join(d, "shim.o"), os.path.join(d, "oracle.o"), "-lm"], check=True)
    L = C.CDLL(lib)
    L.mpmc_host_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
    return L


def _run(L, inp, P, steps):
    log = np.zeros((steps, 5))
    summary = np.zeros(8)
    n = C.c_int()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(inp))
    try:
        rc = L.mpmc_host_run(os.path.basename(inp).encode(), P, steps, log.ctypes.data_as(C.c_void_p), steps, C.byref(n), summary.ctypes.data_as(C.c_void_p))
    finally:
        os.chdir(cwd)
    assert rc == 0, rc
    return log[: n.value], summary


# steps replayed on the CPU (the oracle is O(N^2) per energy): enough to go through every move type many times
STEPS = {"traj_nvt_lj216": 3000, "traj_nvt_kat_gs_ranked": 1500, "traj_uvt_pore": 1500, "traj_pi_argon_dimer": 10000, "traj_pi_h2_27x8": 3000,
         "traj_pi_h2_orient_site1": 3000, "traj_pi_h2_orient_site0": 2000}


@pytest.mark.parametrize("name", sorted(cases.TRAJ))
def test_host_drivers_reproduce_reference_trajectory_on_cpu(host_cpu, name, tmp_path):
    s, r = cases.load_golden_traj(name)
    P = int(r["P"])
    ref = r["traj"][: STEPS[name]]
    inp = W.write_reference_job(s, str(tmp_path))
    log, summary = _run(host_cpu, inp, P, len(ref))
    assert len(log) == len(ref)
    same = (log[:, 0] == ref[:, 0]) & (log[:, 3] == ref[:, 3])
    first_bad = -1 if same.all() else int(np.argmin(same))
    assert first_bad == -1, "trajectory diverges at step %d: ours %s reference %s" % (first_bad, log[first_bad], ref[first_bad])
    fin = np.isfinite(ref[:, 1]) & (np.abs(ref[:, 1]) < 1e30)
    scale = np.maximum(np.abs(ref[fin, 1]), 1.0)
    tol = 1e-10 if s.opts.get("polarization") != "on" else 5e-8     # the classic total cancels ~1e5 K of Ewald sub-terms
    assert (np.abs(log[fin, 1] - ref[fin, 1]) / scale).max() < tol
    bfin = np.isfinite(ref[:, 2])          # the orientational term makes the factor exp(+-1e26): 0 or +inf, which must match as such
    assert (log[~bfin, 2] == ref[~bfin, 2]).all()
    assert (np.abs(log[bfin, 2] - ref[bfin, 2]) / np.maximum(np.abs(ref[bfin, 2]), 1e-300) < 1e-6).all()
    assert summary[6] == ref[:, 3].sum() and summary[7] == len(ref) - ref[:, 3].sum()
    if P:
        assert np.allclose(log[:, 4], ref[:, 4], rtol=1e-10, atol=0)


@pytest.mark.parametrize("name", sorted(cases.SHIPPED))
def test_shipped_sample_directory_runs_unmodified_on_cpu(host_cpu, name, tmp_path):
    """A sample directory shipped with the reference (BASELINE config 1: sample-input/pi000-free-argon-2K, `-P 8 equilibrate.in`), files
    byte for byte as shipped — CRYST1 / BOX pseudo-molecule / CONECT records in the PQR, blank lines in the input file — through the
    mirror's own readers and its path-integral driver: every decision and the kinetic-energy series of the reference
    (src/SimulationControl.PathIntegral.cpp:810-828)."""
    files, inp, P, ref = cases.load_shipped(name)
    cases.write_shipped(files, str(tmp_path))
    log, summary = _run(host_cpu, os.path.join(str(tmp_path), inp), P, len(ref))
    assert len(log) == len(ref)
    same = (log[:, 0] == ref[:, 0]) & (log[:, 3] == ref[:, 3])
    assert same.all(), "trajectory diverges at step %d" % int(np.argmin(same))
    assert np.allclose(log[:, 4], ref[:, 4], rtol=1e-10, atol=0)          # kinetic energy after every step
    assert (np.abs(log[:, 1] - ref[:, 1]) <= 1e-10 * np.maximum(np.abs(ref[:, 1]), 1.0)).all()     # potential (identically 0 for free argon)
    assert (np.abs(log[:, 2] - ref[:, 2]) <= 1e-9 * np.maximum(np.abs(ref[:, 2]), 1e-300)).all()
    assert summary[6] == ref[:, 3].sum()


@pytest.mark.parametrize("name", sorted(cases.MC_AVERAGES))
def test_averages_accumulated_in_the_loop_match_the_reference(host_cpu, name, tmp_path):
    """System::mc of the mirror averages its observables every correlation time and at the end like the reference's loop does
    (do_corrtime_bookkeeping -> update_root_averages, src/System.MonteCarlo.cpp:104-106, src/System.Averages.cpp:8-208): energies,
    N, density, heat capacity, compressibility, weight percent, excess adsorption, pore density, qst against what the unmodified
    reference accumulated over the same seeded chain (tests/golden/mc_averages.npz)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "mc_averages.npz"))
    build, steps, corrtime = cases.MC_AVERAGES[name]
    s = build()
    s.opts.update({"numsteps": str(steps), "corrtime": str(corrtime), "pqr_restart": "off", "pqr_output": "off"})
    inp = W.write_reference_job(s, str(tmp_path))
    log, summary = _run(host_cpu, inp, 0, steps)
    assert len(log) == steps
    host_cpu.mpmc_host_last_averages.argtypes = [C.c_void_p]
    o = np.zeros(27)
    host_cpu.mpmc_host_last_averages(o.ctypes.data_as(C.c_void_p))
    ref = z[name]
    assert o[24] == 1 + steps // corrtime + (1 if steps % corrtime else 0)       # the initial state counts once (setup_mpi)
    assert o[22] == ref[22] and o[23] == ref[23]                          # frozen mass, volume
    for k, a, b in zip(z["keys"][:22], o[:22], ref[:22]):
        k = str(k)
        if np.isnan(b):
            assert np.isnan(a), k                                         # e.g. the density error of a constant-N chain: sqrt of a rounding residue
        elif b == 0.0:
            assert a == 0.0, k
        elif k in ("N", "N_error", "density", "density_error", "percent_wt", "percent_wt_me", "excess_ratio", "pore_density", "compressibility",
                   "compressibility_error") and name.startswith("nvt"):
            assert a == b, (k, a, b)                                      # functions of N alone: the same arithmetic, the same bits
        else:
            tol = 1e-6 if k.endswith("_error") or k in ("heat_capacity", "qst") else 1e-9    # (differences of nearly equal means)
            assert abs(a - b) <= tol * abs(b), (k, a, b)


@pytest.mark.parametrize("name", sorted(cases.PI_AVERAGES))
def test_path_integral_averages_match_the_reference(host_cpu, name, tmp_path):
    """PI_nvt_mc of the mirror averages the aggregate observables — total, kinetic estimator, potential terms, N, density, heat
    capacity, compressibility — for the initial state, every correlation time and at the end, like the reference
    (src/SimulationControl.PathIntegral.cpp:63-66, 176-178, 211-270) — against what the unmodified reference accumulates over the same
    seeded chain (tests/golden/pi_averages.npz)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "pi_averages.npz"))
    build, P, steps, corrtime = cases.PI_AVERAGES[name]
    s = build()
    s.opts.update({"numsteps": str(steps), "corrtime": str(corrtime), "pqr_restart": "off", "pqr_output": "off"})
    inp = W.write_reference_job(s, str(tmp_path))
    log, summary = _run(host_cpu, inp, P, steps)
    assert len(log) == steps
    host_cpu.mpmc_host_last_averages.argtypes = [C.c_void_p]
    o = np.zeros(27)
    host_cpu.mpmc_host_last_averages(o.ctypes.data_as(C.c_void_p))
    ours = {"energy": o[0], "energy_error": o[1], "N": o[2], "N_error": o[3], "coulombic_energy": o[4], "coulombic_energy_error": o[5],
            "rd_energy": o[6], "rd_energy_error": o[7], "polarization_energy": o[8], "polarization_energy_error": o[9], "density": o[10],
            "density_error": o[11], "heat_capacity": o[12], "heat_capacity_error": o[13], "compressibility": o[14], "compressibility_error": o[15],
            "frozen_mass": o[22], "volume": o[23], "kinetic_energy": o[25], "kinetic_energy_error": o[26]}
    assert o[24] == 1 + steps // corrtime + (1 if steps % corrtime else 0)       # the initial state counts once
    for k, b in zip(z["keys"], z[name]):
        k, a = str(k), ours[str(k)]
        if np.isnan(b):
            assert np.isnan(a), k
        elif b == 0.0:
            assert a == 0.0, k
        elif k in ("N", "N_error", "density", "density_error", "compressibility", "compressibility_error", "frozen_mass", "volume"):
            assert a == b, (k, a, b)                                              # functions of N and the cell alone: the same bits
        else:
            tol = 1e-6 if k.endswith("_error") or k == "heat_capacity" else 1e-9
            assert abs(a - b) <= tol * abs(b), (k, a, b)


@pytest.mark.parametrize("name", sorted(cases.FINAL_PQR))
def test_final_state_file_after_a_run_matches_the_reference(host_cpu, name, tmp_path):
    """The geometry a chain ends with — after displacements, insertions and removals (uVT) or bead moves (path integrals) — written by
    the mirror as its final PQR file (`pqr_output`; `-000k` per bead system) is byte for byte the file the reference writes for its
    own state after the same seeded chain (tests/golden/final_pqr.npz)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "final_pqr.npz"))
    build, P, steps, sidx = cases.FINAL_PQR[name]
    s = build()
    s.opts.update({"numsteps": str(steps), "corrtime": "1000000", "pqr_restart": "off", "pqr_output": "final_state.pqr"})
    inp = W.write_reference_job(s, str(tmp_path))
    log, summary = _run(host_cpu, inp, P, steps)
    assert len(log) == steps
    fn = "final_state.pqr" if not P else "final_state-%04d.pqr" % sidx
    text = open(os.path.join(str(tmp_path), fn)).read()
    want = str(z[name])
    if text != want:
        a, b = text.splitlines(), want.splitlines()
        bad = [i for i in range(min(len(a), len(b))) if a[i] != b[i]]
        raise AssertionError("%d / %d lines differ (lengths %d, %d); first: %r vs %r" % (len(bad), len(b), len(a), len(b),
                             a[bad[0]] if bad else None, b[bad[0]] if bad else None))
