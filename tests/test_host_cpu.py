"""CPU-side checks of the C++ host mirror: it builds and loads, parses the reference's input keywords, and fails with the
reference's own error codes (src/constants.h:108-147) before any device work; with no GPU the energy call fails loudly."""
import os

import pytest

from mpmcxx_b200 import host_binding, workloads as W
from tests import cases


def _job(tmp_path, system, extra=None, drop=()):
    s = system.copy()
    s.opts.update({"seed": "1", "numsteps": "5"})
    s.opts.update(extra or {})
    for k in drop:
        s.opts.pop(k, None)
    return W.write_reference_job(s, str(tmp_path))


def _code(fn):
    with pytest.raises(RuntimeError) as ei:
        fn()
    return int(str(ei.value).split()[-1])


def test_parser_error_codes(tmp_path):
    lj = W.lj_lattice(3, 20.0)
    inp = _job(tmp_path / "a", lj, {"no_such_keyword": "1"})
    assert _code(lambda: host_binding.run(inp)) == 3000              # invalid_input
    inp = _job(tmp_path / "b", lj, drop=("temperature",))
    with open(inp) as f:
        text = "".join(l for l in f if not l.startswith("temperature"))
    open(inp, "w").write(text)
    assert _code(lambda: host_binding.run(inp)) == 3000              # failed validation -> invalid_input, like the reference's constructor
    inp = _job(tmp_path / "c", lj, {"feynman_hibbs": "on"})
    assert _code(lambda: host_binding.run(inp)) == 4004              # unsupported_setting
    tmpl, _ = W.pi_h2_cluster(n_side=2, P=8, L=30.0)
    inp = _job(tmp_path / "d", tmpl)
    assert _code(lambda: host_binding.run(inp, P=6)) == 12000        # Trotter number must be a power of two >= 4
    inp = _job(tmp_path / "e", tmpl, {"PI_trial_chain_length": "8"})
    assert _code(lambda: host_binding.run(inp, P=8)) == 4000         # chain length must be < P


def test_no_gpu_means_loud_failure(tmp_path):
    import ctypes as C
    from mpmcxx_b200 import engine
    n = C.c_int(0)
    have_gpu = engine.lib().mpmc_device_count(C.byref(n)) == 0 and n.value > 0
    inp = _job(tmp_path, W.lj_lattice(3, 20.0))
    if have_gpu:
        assert host_binding.energy(inp)["rd"] < 0
    else:
        assert _code(lambda: host_binding.energy(inp)) == 30000       # MPMC_ERR_CUDA: no CPU fallback
    assert os.path.exists(os.path.join(os.path.dirname(host_binding.__file__), "mpmcxx-b200"))


@pytest.mark.parametrize("name", ["lj_lattice_4", "tri_gs_ranked_palmo", "h2fw_6_jacobi10", "tri_alpha_set", "pi_h2_five_8x4", "tri_box_from_pqr"])
def test_readers_match_the_reference(tmp_path, name):
    """The mirror's input-file and PQR readers against the reference's own (src/SimulationControl.cpp:1090-2300,
    src/System.cpp:361-700, read by oracle/ref_harness.cpp into tests/golden/parsed_*.npz): same flat site table in the same order
    — coordinates, charges in the reference's internal units, polarizabilities, LJ parameters, masses, molecule index, frozen flag —
    and the same cell (basis, reciprocal basis, volume, cutoff, Ewald alphas).  No GPU involved."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "parsed_%s.npz" % name))
    d = tmp_path / name
    d.mkdir()
    (d / "input.in").write_text(str(g["input_in"]))
    (d / "input.pqr").write_text(str(g["input_pqr"]))
    got = host_binding.describe(str(d / "input.in"), P=int(g["P"]))
    for k in ("pos", "charge", "alpha", "eps", "sigma", "mass"):
        assert got[k].shape == g["ref_" + k].shape, k
        assert np.array_equal(got[k], g["ref_" + k]), (k, np.abs(got[k] - g["ref_" + k]).max())
    assert np.array_equal(got["mol"], g["ref_mol"]) and np.array_equal(got["frozen"] != 0, g["ref_frozen"] != 0)
    for k in ("basis", "recip"):
        assert np.array_equal(got[k], g["cell_" + k]), k
    for k in ("volume", "cutoff", "ewald_alpha", "polar_ewald_alpha"):
        assert got[k] == float(g["cell_" + k]), (k, got[k], float(g["cell_" + k]))


def test_malformed_jobs_are_refused_with_the_reference_error_codes(tmp_path):
    """The mirror's reader and validator throw what the reference throws for the same malformed job (its constructor turns a failed
    validation into invalid_input, src/SimulationControl.cpp:67-72; the path-integral checks throw their own codes, :552-606);
    codes recorded from the unmodified reference by tests/golden/make_golden.py errors."""
    import json
    from mpmcxx_b200 import host_binding
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_errors.json")))
    assert set(golden) == set(cases.INPUT_ERRORS)
    for name, (build, P, mutate) in cases.INPUT_ERRORS.items():
        s = build()
        mutate(s)
        d = tmp_path / name
        d.mkdir()
        inp = W.write_reference_job(s, str(d))
        try:
            host_binding.describe(inp, P=P)
            code = 0
        except RuntimeError as e:
            code = int(str(e).split()[-1])
        assert code == golden[name], (name, code, golden[name])
