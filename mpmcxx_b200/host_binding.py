"""ctypes binding of libmpmc_host.so (the C++ mirror of System/SimulationControl over the engine), for the trajectory tests."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build_host()
        L = C.CDLL(path)
        L.mpmc_host_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        L.mpmc_host_run_sharded.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        L.mpmc_host_write_pqr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        L.mpmc_host_last_stats.argtypes = [C.c_void_p]
        L.mpmc_host_last_averages.argtypes = [C.c_void_p]
        L.mpmc_host_last_stats.restype = None
        L.mpmc_host_energy.argtypes = [C.c_char_p, C.c_void_p]
        L.mpmc_host_describe.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)] + [C.c_void_p] * 9
        _lib = L
    return _lib


def run(input_file: str, P: int = 0, max_steps: int = 0, capacity: int = 20000):
    """Run the simulation described by `input_file` from its directory.  Returns (log[n,5], summary[8])."""
    log = np.zeros((capacity, 5))
    summary = np.zeros(8)
    n = C.c_int()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(input_file)))
    try:
        rc = lib().mpmc_host_run(os.path.basename(input_file).encode(), P, max_steps, log.ctypes.data_as(C.c_void_p), capacity, C.byref(n),
                                 summary.ctypes.data_as(C.c_void_p))
    finally:
        os.chdir(cwd)
    if rc:
        raise RuntimeError("host run failed with code %d" % rc)
    return log[: n.value].copy(), summary


def run_sharded(input_file: str, P: int, rank: int, nranks: int, device: int, nccl_id: bytes, max_steps: int = 0, capacity: int = 20000):
    """A path-integral run with the bead systems sharded over `nranks` processes (one GPU each); call from every rank with the same
    128-byte id (mpmcxx_b200.engine.nccl_unique_id() of rank 0).  Returns (log[n,5], summary[8]), identical on every rank."""
    log = np.zeros((capacity, 5))
    summary = np.zeros(8)
    n = C.c_int()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(input_file)))
    try:
        rc = lib().mpmc_host_run_sharded(os.path.basename(input_file).encode(), P, max_steps, rank, nranks, device, nccl_id,
                                         log.ctypes.data_as(C.c_void_p), capacity, C.byref(n), summary.ctypes.data_as(C.c_void_p))
    finally:
        os.chdir(cwd)
    if rc:
        raise RuntimeError("sharded host run failed with code %d" % rc)
    return log[: n.value].copy(), summary


def write_pqr(input_file: str, out_path: str = "", P: int = 0, s: int = 0):
    """The PQR file the mirror writes for (bead system s of) the job, as read (no GPU needed), and the file names it chose for that
    system: returns (pqr_input, pqr_restart, pqr_output)."""
    names = C.create_string_buffer(4096)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(input_file)))
    try:
        rc = lib().mpmc_host_write_pqr(os.path.basename(input_file).encode(), P, s, out_path.encode(), names, 4096)
    finally:
        os.chdir(cwd)
    if rc:
        raise RuntimeError("host write_pqr failed with code %d" % rc)
    return tuple(names.value.decode().split("\n"))


AVERAGE_KEYS = ("energy", "energy_error", "N", "N_error", "coulombic_energy", "coulombic_energy_error", "rd_energy", "rd_energy_error",
                "polarization_energy", "polarization_energy_error", "density", "density_error", "heat_capacity", "heat_capacity_error",
                "compressibility", "compressibility_error", "percent_wt", "percent_wt_me", "excess_ratio", "qst", "pore_density", "NU",
                "frozen_mass", "volume", "samples", "kinetic_energy", "kinetic_energy_error")


def last_averages():
    """What the last classic run (nvt / uvt) averaged every correlation time and at its end, the way the reference does
    (System::update_root_averages, src/System.Averages.cpp:8-208)."""
    o = np.zeros(27)
    lib().mpmc_host_last_averages(o.ctypes.data_as(C.c_void_p))
    return dict(zip(AVERAGE_KEYS, o.tolist()))


def last_stats():
    """(wall seconds of the last run's step loop, potential sweeps made in it)"""
    o = np.zeros(2)
    lib().mpmc_host_last_stats(o.ctypes.data_as(C.c_void_p))
    return float(o[0]), float(o[1])


def energy(input_file: str):
    out = np.zeros(5)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(input_file)))
    try:
        rc = lib().mpmc_host_energy(os.path.basename(input_file).encode(), out.ctypes.data_as(C.c_void_p))
    finally:
        os.chdir(cwd)
    if rc:
        raise RuntimeError("host energy failed with code %d" % rc)
    return dict(energy=out[0], rd=out[1], coulombic=out[2], polar=out[3], iterations=int(out[4]))


def describe(input_file: str, P: int = 0, capacity: int = 200000):
    """What the mirror's readers make of an input file + PQR (no GPU needed): the flat site table and the cell."""
    pos = np.zeros(3 * capacity); q = np.zeros(capacity); al = np.zeros(capacity); ep = np.zeros(capacity); sg = np.zeros(capacity)
    ms = np.zeros(capacity); mol = np.zeros(capacity, np.int32); fz = np.zeros(capacity, np.int32); cell = np.zeros(22)
    n = C.c_int()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(input_file)))
    try:
        rc = lib().mpmc_host_describe(os.path.basename(input_file).encode(), P, capacity, C.byref(n),
                                      *[a.ctypes.data_as(C.c_void_p) for a in (pos, q, al, ep, sg, ms, mol, fz, cell)])
    finally:
        os.chdir(cwd)
    if rc:
        raise RuntimeError("host describe failed with code %d" % rc)
    k = n.value
    return dict(pos=pos[:3 * k].reshape(k, 3).copy(), charge=q[:k].copy(), alpha=al[:k].copy(), eps=ep[:k].copy(), sigma=sg[:k].copy(),
                mass=ms[:k].copy(), mol=mol[:k].copy(), frozen=fz[:k].copy(), basis=cell[:9].reshape(3, 3).copy(),
                recip=cell[9:18].reshape(3, 3).copy(), volume=cell[18], cutoff=cell[19], ewald_alpha=cell[20], polar_ewald_alpha=cell[21])
