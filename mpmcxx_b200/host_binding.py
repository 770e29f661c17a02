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
        L.mpmc_host_energy.argtypes = [C.c_char_p, C.c_void_p]
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
