"""The host mirror's file formats against the reference's own code, without a GPU: the PQR WRITER byte for byte
(System::write_molecules, src/System.Output.cpp:900-1091: CRYST1, PDB-style or extended coordinates, wrapped or not, the BOX
pseudo-molecule with its CONECT records, the basis remarks), the per-system file names (Output::make_filename, src/Output.cpp:46-92;
check_io_files_options, src/SimulationControl.cpp:2196-2360) and the RESTART SELECTION of a `parallel_restarts on` job — the reference's
shipped sample-input/pi000-free-argon-2K/input.in with the restart files that ship beside it, unmodified.  Golden data:
tests/golden/pqr_written.npz, written by the unmodified reference (tests/golden/make_golden.py written)."""
import os

import numpy as np
import pytest

from mpmcxx_b200 import host_binding, workloads as W
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

Z = np.load(os.path.join(cases.GOLDEN_DIR, "pqr_written.npz"))


@pytest.mark.parametrize("name", sorted(cases.WRITTEN))
def test_pqr_writer_matches_the_reference_byte_for_byte(name, tmp_path):
    build, P, sidx = cases.WRITTEN[name]
    inp = W.write_reference_job(build(), str(tmp_path))
    out = str(tmp_path / "written.pqr")
    names = host_binding.write_pqr(inp, out, P=P, s=max(sidx, 0))
    assert open(out).read() == str(Z["text_" + name])
    assert "\n".join(names) == str(Z["names_" + name])
    # writing again moves the previous file to "<name>.last" (write_molecules_wrapper, :883-888)
    host_binding.write_pqr(inp, out, P=P, s=max(sidx, 0))
    assert open(out + ".last").read() == str(Z["text_" + name])


def test_written_pqr_reads_back_to_the_same_sites(tmp_path):
    """reader(writer(system)) == system to the writer's printed precision: what a restart relies on."""
    s = W.triclinic_mix(solver=W.SOLVER_GS_RANKED_PALMO)
    s.opts["wrapall"] = "off"
    inp = W.write_reference_job(s, str(tmp_path))
    out = str(tmp_path / "round.pqr")
    host_binding.write_pqr(inp, out)
    d0 = host_binding.describe(inp)
    s2 = s.copy()
    s2.opts["pqr_input"] = "round.pqr"
    W.write_input(s2, str(tmp_path / "again.in"), "round.pqr")
    d1 = host_binding.describe(str(tmp_path / "again.in"))
    assert np.abs(d1["pos"] - d0["pos"]).max() <= 0.5e-3 + 1e-12          # %8.3f
    assert np.allclose(d1["charge"], d0["charge"], rtol=0, atol=0.5e-5 * W.E2REDUCED + 1e-9)
    assert np.array_equal(d1["mol"], d0["mol"]) and np.array_equal(d1["frozen"], d0["frozen"])


def test_restart_selection_of_the_shipped_parallel_restart_job(tmp_path):
    """sample-input/pi000-free-argon-2K/input.in as shipped (`parallel_restarts on`), `-P 8`: bead systems 0..3 restart from
    Ar2K.restart-000j.pqr, 4..7 from system 0's ".last" copy (the reference builds that name from its rank, 0 in one process)."""
    for f, text in zip(Z["restart_files"], Z["restart_texts"]):
        with open(tmp_path / str(f), "w") as fp:
            fp.write(str(text))
    inp = str(tmp_path / "input.in")
    for i in range(8):
        names = host_binding.write_pqr(inp, "", P=8, s=i)
        assert "\n".join(os.path.basename(x) for x in names) == str(Z["restart_names"][i])
    d = host_binding.describe(inp, P=8)                # the first bead system's sites as read
    assert np.array_equal(d["pos"], Z["restart_pos"][0])


# ---- orientational degree of freedom of path-integral sorbates (src/Molecule.cpp:211-254, src/SimulationControl.cpp:306-339) ----
def test_orient_turns_the_handle_site_onto_the_orientation_and_nothing_else(tmp_path):
    import ctypes as C
    import subprocess
    lib = str(tmp_path / "libhost_orient.so")
    host = os.path.join(ROOT, "mpmcxx_b200", "host")
    # Molecule lives in host.cpp, which also holds System::energy(): link it against the oracle-backed shim like the trajectory tests
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle")]
    cflags = ["-O2", "-std=c11", "-fPIC", "-fopenmp", "-ffp-contract=off"]
    subprocess.run(["gcc"] + cflags + inc + ["-c", os.path.join(ROOT, "oracle", "oracle.c"), "-o", str(tmp_path / "oracle.o")], check=True)
    subprocess.run(["gcc"] + cflags + inc + ["-c", os.path.join(ROOT, "tests", "shim", "oracle_engine.c"), "-o", str(tmp_path / "shim.o")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp"] + inc + ["-o", lib] +
                   [os.path.join(host, f) for f in ("host.cpp", "sim_control.cpp", "host_capi.cpp")] + [str(tmp_path / "shim.o"), str(tmp_path / "oracle.o"), "-lm"], check=True)
    L = C.CDLL(lib)
    L.mpmc_host_debug_orient.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rs = np.random.RandomState(5)
    # a five-site linear molecule (centre site on the COM, +-0.371, +-0.329 along a random axis), somewhere in space
    axis = rs.normal(size=3); axis /= np.linalg.norm(axis)
    off = np.array([0.0, 0.371, -0.371, 0.329, -0.329])
    mass = np.array([0.0, 1.008, 1.008, 0.0, 0.0])
    centre = np.array([3.0, -7.0, 11.0])
    pos0 = centre + off[:, None] * axis
    for site, o in ((1, rs.normal(size=3)), (2, np.array([0.0, 0.0, 2.5])), (3, rs.normal(size=3) * 0.3)):   # (generic directions: an orientation exactly
    # opposite to the handle leaves the reference's rotation axis, a cross product, to rounding noise)
        pos = np.ascontiguousarray(pos0.copy())
        o = np.ascontiguousarray(o, dtype=np.float64)
        assert L.mpmc_host_debug_orient(5, pos.ctypes.data_as(C.c_void_p), mass.ctypes.data_as(C.c_void_p), site, o.ctypes.data_as(C.c_void_p)) == 0
        com = (mass[:, None] * pos).sum(0) / mass.sum()
        assert np.allclose(com, centre, atol=1e-12)                                       # the COM stays
        d0 = np.linalg.norm(pos0[:, None] - pos0[None], axis=-1)
        assert np.allclose(np.linalg.norm(pos[:, None] - pos[None], axis=-1), d0, atol=1e-12)      # rigid
        h = pos[site] - com
        assert np.allclose(h / np.linalg.norm(h), o / np.linalg.norm(o), atol=1e-9)              # the handle points along the orientation
    # the reference's usual case: the handle is the site ON the centre of mass -> no direction, identity rotation
    pos = np.ascontiguousarray(pos0.copy())
    o = np.array([0.3, -0.2, 0.9])
    assert L.mpmc_host_debug_orient(5, pos.ctypes.data_as(C.c_void_p), mass.ctypes.data_as(C.c_void_p), 0, o.ctypes.data_as(C.c_void_p)) == 0
    assert np.allclose(pos, pos0, atol=1e-12)


def test_malformed_orientation_keyword_is_rejected(tmp_path):
    """sorbate_orientation_site / sorbate_bondlength / sorbate_reducedMass take `<type> <value>`; a malformed line is rejected like
    any other (invalid_input, 3000)."""
    from mpmcxx_b200 import host_binding
    tmpl, _ = W.pi_h2_cluster(n_side=2, P=4, L=30.0, five_site=True)
    tmpl.opts.update({"seed": "1", "numsteps": "2", "PI_trial_chain_length": "1", "_lines": ["sorbate_bondlength H2"]})
    inp = W.write_reference_job(tmpl, str(tmp_path))
    L = host_binding.lib()
    import ctypes as C
    n = C.c_int()
    rc = L.mpmc_host_describe(inp.encode(), 4, 0, C.byref(n), *([None] * 9))
    assert rc == 3000
