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
