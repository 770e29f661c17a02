"""ctypes binding of oracle/liboracle.so (oracle.c, our C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

IOPT = ["rd_lrc", "rd_only", "polarization", "damp_type", "polar_ewald", "polar_iterative", "polar_gs", "polar_gs_ranked",
        "polar_palmo", "polar_sor", "polar_esor", "polar_zodid", "polar_rrms", "polar_max_iter", "ewald_kmax"]
DOPT = ["polar_damp", "polar_gamma", "polar_precision", "ewald_alpha", "polar_ewald_alpha"]
OUT = ["energy", "rd_total", "coulombic", "polar", "rd", "lrc_pair", "lrc_self", "es_real", "es_self_intra", "es_recip",
       "es_self", "iterations", "dipole_rrms", "iterator_failed", "volume", "cutoff", "ewald_alpha", "polar_ewald_alpha",
       "n_kvec", "n_pairs_in_cutoff", "rd_abs", "es_real_abs", "es_intra_abs"]


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.orc_energy.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, _ip, _dp, _dp, vp, vp, vp, vp, vp]
        L.orc_pi_energy.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, _ip, _dp, C.c_double, _dp, vp]
        L.orc_cell.argtypes = [_dp, C.c_double, C.c_double, _dp]
        _lib = L
    return _lib


def _opts(system):
    from mpmcxx_b200 import config
    o = config.from_keywords(system.opts).as_dict()
    iopt = np.zeros(16, np.int32)
    dopt = np.zeros(8, np.float64)
    for i, k in enumerate(IOPT):
        iopt[i] = o[k]
    for i, k in enumerate(DOPT):
        dopt[i] = o[k]
    return iopt, dopt, o


def cell(basis, ewald_alpha=0.0, polar_ewald_alpha=0.0):
    o = np.zeros(22)
    lib().orc_cell(np.ascontiguousarray(basis, np.float64).reshape(-1), ewald_alpha, polar_ewald_alpha, o)
    return dict(basis=o[:9].reshape(3, 3).copy(), recip=o[9:18].reshape(3, 3).copy(), volume=o[18], cutoff=o[19],
                ewald_alpha=o[20], polar_ewald_alpha=o[21])


def energy(system, pos=None, want_sites: bool = True):
    """Cold System::energy() of a SiteSystem -> dict of components (+ per-site mu / fields when polarization is on)."""
    iopt, dopt, _ = _opts(system)
    n = system.n
    p = np.ascontiguousarray(system.pos if pos is None else pos, np.float64).reshape(-1)
    out = np.zeros(24)
    arrs = [np.zeros(3 * n) for _ in range(4)] + [np.zeros(n)] if want_sites else [None] * 5
    ptrs = [a.ctypes.data_as(C.c_void_p) if a is not None else None for a in arrs]
    lib().orc_energy(n, p, np.ascontiguousarray(system.charge), system.alpha, system.eps, system.sigma, system.mol, system.frozen,
                     np.ascontiguousarray(system.basis).reshape(-1), iopt, dopt, out, *ptrs)
    res = {k: float(out[i]) for i, k in enumerate(OUT)}
    if want_sites:
        res.update(mu=arrs[0].reshape(n, 3), ef_static=arrs[1].reshape(n, 3), ef_induced=arrs[2].reshape(n, 3),
                   ef_induced_change=arrs[3].reshape(n, 3), rank_metric=arrs[4])
    return res


def pi_energy(system, beads, temperature=None):
    iopt, dopt, o = _opts(system)
    P, n = beads.shape[0], system.n
    T = o["temperature"] if temperature is None else temperature
    out = np.zeros(8)
    per = np.zeros(4 * P)
    lib().orc_pi_energy(P, n, np.ascontiguousarray(beads, np.float64).reshape(-1), np.ascontiguousarray(system.charge),
                        system.alpha, system.eps, system.sigma, system.mass, system.mol, system.frozen,
                        np.ascontiguousarray(system.basis).reshape(-1), iopt, dopt, float(T), out, per.ctypes.data_as(C.c_void_p))
    return dict(potential=out[0], rd=out[1], coulombic=out[2], polar=out[3], vdw=out[4], chain_mass_len2=out[5],
                kinetic=out[6], per_bead=per.reshape(P, 4))
