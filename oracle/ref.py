"""ctypes binding of oracle/_ref/libmpmc_ref.so (the unmodified reference + ref_harness.cpp).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmpmc_ref.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.mref_open.restype = C.c_void_p
        L.mref_open.argtypes = [C.c_char_p, C.c_int]
        L.mref_last_error.restype = C.c_int
        L.mref_nsys.argtypes = [C.c_void_p]
        L.mref_natoms.argtypes = [C.c_void_p, C.c_int]
        L.mref_get_sites.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
        L.mref_set_pos.argtypes = [C.c_void_p, C.c_int, _dp]
        L.mref_get_cell.argtypes = [C.c_void_p, C.c_int, _dp]
        L.mref_energy.argtypes = [C.c_void_p, C.c_int, _dp]
        L.mref_terms.argtypes = [C.c_void_p, C.c_int, _dp]
        L.mref_get_dipoles.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.mref_pi_energy.argtypes = [C.c_void_p, _dp]
        L.mref_mc_trajectory.argtypes = [C.c_void_p, C.c_int, _dp]
        L.mref_pi_trajectory.argtypes = [C.c_void_p, C.c_int, _dp]
        L.mref_write_pqr.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.mref_io_filenames.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.mref_root_averages.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.mref_mc_averages.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mref_pi_averages.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mref_pi_potential.argtypes = [C.c_void_p]
        L.mref_pi_potential.restype = C.c_double
        _lib = L
    return _lib


TERM_KEYS = ["lj_total", "rd", "lrc_pair", "lrc_self", "es_real_minus_intra", "es_real", "es_self_intra", "es_recip",
             "es_self", "polar", "iterations", "dipole_rrms", "iterator_failed", "n_pairs", "n_pairs_in_cutoff", "_"]


class RefSystem:
    """One reference SimulationControl built from a SiteSystem (written to a temp dir as input.in + input.pqr)."""

    def __init__(self, system, P: int = 0, workdir: str | None = None, ensemble: str | None = None):
        from mpmcxx_b200 import workloads
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="mref_")
            workdir = self._tmp.name
        s = system.copy()
        if ensemble is not None:
            s.opts["ensemble"] = ensemble
            if ensemble != "uvt":
                for k in ("pressure", "insert_probability", "h2_fugacity", "free_volume"):
                    s.opts.pop(k, None)
        inp = workloads.write_reference_job(s, workdir)
        cwd = os.getcwd()
        os.chdir(workdir)          # the reference opens pqr_input relative to cwd
        try:
            self.h = lib().mref_open(inp.encode(), P)
        finally:
            os.chdir(cwd)
        if not self.h:
            raise RuntimeError("reference refused the input: error code %d" % lib().mref_last_error())
        self.P = lib().mref_nsys(self.h)

    @classmethod
    def from_directory(cls, directory: str, input_name: str, P: int = 0):
        """A reference SimulationControl opened on an existing job directory exactly as it is (e.g. a copy of one of the
        reference's shipped sample-input directories)."""
        self = cls.__new__(cls)
        self._tmp = None
        cwd = os.getcwd()
        os.chdir(directory)
        try:
            self.h = lib().mref_open(input_name.encode(), P)
        finally:
            os.chdir(cwd)
        if not self.h:
            raise RuntimeError("reference refused the input: error code %d" % lib().mref_last_error())
        self.P = lib().mref_nsys(self.h)
        return self

    def natoms(self, s: int = -1) -> int:
        return lib().mref_natoms(self.h, s)

    def sites(self, s: int = -1):
        n = self.natoms(s)
        pos = np.zeros(3 * n); q = np.zeros(n); al = np.zeros(n); ep = np.zeros(n); sg = np.zeros(n); ms = np.zeros(n)
        mol = np.zeros(n, np.int32); fz = np.zeros(n, np.int32)
        lib().mref_get_sites(self.h, s, pos, q, al, ep, sg, ms, mol, fz)
        return dict(pos=pos.reshape(n, 3), charge=q, alpha=al, eps=ep, sigma=sg, mass=ms, mol=mol, frozen=fz)

    def set_pos(self, pos, s: int = -1) -> None:
        lib().mref_set_pos(self.h, s, np.ascontiguousarray(pos, dtype=np.float64).reshape(-1))

    def cell(self, s: int = -1):
        o = np.zeros(22)
        lib().mref_get_cell(self.h, s, o)
        return dict(basis=o[:9].reshape(3, 3).copy(), recip=o[9:18].reshape(3, 3).copy(), volume=o[18], cutoff=o[19],
                    ewald_alpha=o[20], polar_ewald_alpha=o[21])

    def energy(self, s: int = -1):
        o = np.zeros(16)
        rc = lib().mref_energy(self.h, s, o)
        if rc:
            raise RuntimeError("reference energy() threw %d" % rc)
        return dict(energy=o[0], rd=o[1], coulombic=o[2], polar=o[3], vdw=o[4], iterations=o[5], dipole_rrms=o[6],
                    iterator_failed=int(o[7]), N=o[8])

    def terms(self, s: int = -1):
        o = np.zeros(16)
        rc = lib().mref_terms(self.h, s, o)
        if rc:
            raise RuntimeError("reference term function threw %d" % rc)
        return {k: float(v) for k, v in zip(TERM_KEYS, o) if k != "_"}

    def dipoles(self, s: int = -1):
        n = self.natoms(s)
        mu = np.zeros(3 * n); efs = np.zeros(3 * n); efi = np.zeros(3 * n); efc = np.zeros(3 * n); rk = np.zeros(n)
        lib().mref_get_dipoles(self.h, s, mu, efs, efi, efc, rk)
        return dict(mu=mu.reshape(n, 3), ef_static=efs.reshape(n, 3), ef_induced=efi.reshape(n, 3),
                    ef_induced_change=efc.reshape(n, 3), rank_metric=rk)

    def pi_energy(self):
        o = np.zeros(8)
        rc = lib().mref_pi_energy(self.h, o)
        if rc:
            raise RuntimeError("reference PI energy threw %d" % rc)
        return dict(potential=o[0], rd=o[1], coulombic=o[2], polar=o[3], vdw=o[4], kinetic=o[5], chain_mass_len2=o[6])

    def mc_trajectory(self, nsteps: int):
        log = np.zeros(5 * nsteps)
        rc = lib().mref_mc_trajectory(self.h, nsteps, log)
        if rc:
            raise RuntimeError("reference mc loop threw %d" % rc)
        return log.reshape(nsteps, 5)

    def pi_trajectory(self, nsteps: int):
        log = np.zeros(5 * nsteps)
        rc = lib().mref_pi_trajectory(self.h, nsteps, log)
        if rc:
            raise RuntimeError("reference PI loop threw %d" % rc)
        return log.reshape(nsteps, 5)

    def mc_averages(self, nsteps: int, corrtime: int):
        """The classic Markov chain with update_root_averages every `corrtime` steps and at the end -> 25 numbers (see the harness);
        once per process (the reference counts its samples in a function-static)."""
        o = np.zeros(25)
        rc = lib().mref_mc_averages(self.h, nsteps, corrtime, o.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError("reference mc loop threw %d" % rc)
        return o

    def pi_averages(self, nsteps: int, corrtime: int):
        """The path-integral chain with the reference's averaging (initial state + every `corrtime` steps + the end) -> 20 numbers
        (see the harness); once per process."""
        o = np.zeros(20)
        rc = lib().mref_pi_averages(self.h, nsteps, corrtime, o.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError("reference PI loop threw %d" % rc)
        return o

    def root_averages(self, samples, s: int = -1):
        """update_root_averages over `samples` [n, 6] = (energy, coulombic, rd, polarization, N, NU) -> 25 numbers (see the harness).
        The reference counts its samples in a function-static: call once per process."""
        x = np.ascontiguousarray(samples, dtype=np.float64)
        o = np.zeros(25)
        rc = lib().mref_root_averages(self.h, s, len(x), x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError("reference update_root_averages threw %d" % rc)
        return o

    def write_pqr(self, path: str, s: int = -1, after_energy: bool = False) -> None:
        """The reference's own PQR writer (System::write_molecules_wrapper) for system s."""
        rc = lib().mref_write_pqr(self.h, s, path.encode(), 1 if after_energy else 0)
        if rc:
            raise RuntimeError("reference write_molecules_wrapper returned %d" % rc)

    def io_filenames(self, s: int = -1):
        """(pqr_input, pqr_restart, pqr_output) the reference chose for system s."""
        buf = C.create_string_buffer(4096)
        lib().mref_io_filenames(self.h, s, buf, 4096)
        return tuple(buf.value.decode().split("\n"))

    def pi_potential(self) -> float:
        return lib().mref_pi_potential(self.h)
