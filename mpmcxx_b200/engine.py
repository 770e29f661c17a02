"""ctypes binding of the engine's C-ABI (include/mpmc_b200.h) for tests and bench.py.

This is plumbing: numpy host buffers in, dicts of doubles out.  The product is the shared library; if it (or a CUDA
device) is missing this module raises — there is no CPU path to fall back to.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from . import config as _config

_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class MpmcConfig(C.Structure):
    _fields_ = [("basis", C.c_double * 9), ("n_beads", C.c_int), ("capacity", C.c_int), ("device", C.c_int),
                ("rd_lrc", C.c_int), ("rd_only", C.c_int), ("ewald_kmax", C.c_int), ("ewald_alpha", C.c_double),
                ("polarization", C.c_int), ("polar_ewald", C.c_int), ("polar_iterative", C.c_int), ("damp_type", C.c_int),
                ("polar_gs", C.c_int), ("polar_gs_ranked", C.c_int), ("polar_palmo", C.c_int), ("polar_sor", C.c_int),
                ("polar_esor", C.c_int), ("polar_zodid", C.c_int), ("polar_rrms", C.c_int), ("polar_max_iter", C.c_int),
                ("polar_damp", C.c_double), ("polar_gamma", C.c_double), ("polar_precision", C.c_double),
                ("polar_ewald_alpha", C.c_double), ("reserved", C.c_int * 8)]


class MpmcEnergyOut(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("energy", "rd_energy", "coulombic_energy", "polarization_energy", "vdw_energy", "rd_pair",
                                          "rd_lrc_pair", "rd_lrc_self", "es_real", "es_self_intra", "es_reciprocal", "es_self",
                                          "dipole_rrms", "n_pairs_in_cutoff", "n_pair_evals")] + \
               [("polarization_iterations", C.c_int), ("iterator_failed", C.c_int), ("reserved", C.c_int * 4)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class MpmcError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("mpmc error %d: %s" % (code, text))
        self.code = code


EXPORTS = ["mpmc_abi_version", "mpmc_last_error", "mpmc_device_count", "mpmc_create", "mpmc_destroy", "mpmc_set_cell", "mpmc_get_cell",
           "mpmc_upload_sites", "mpmc_update_sites", "mpmc_update_sites_all_beads", "mpmc_insert_sites", "mpmc_remove_sites",
           "mpmc_num_sites", "mpmc_energy", "mpmc_energy_enqueue", "mpmc_energy_fetch", "mpmc_download_dipoles",
           "mpmc_download_rank_metric", "mpmc_pi_potential", "mpmc_pi_chain", "mpmc_nccl_get_unique_id", "mpmc_nccl_init", "mpmc_pi_potential_allreduce",
           "mpmc_pi_chain_allreduce", "mpmc_set_timing", "mpmc_get_timing", "mpmc_debug_gs_profile", "mpmc_stream", "mpmc_kernel_launches",
           "mpmc_probe_fp64_peak", "mpmc_debug_radial_table", "mpmc_debug_cutoff_thresholds", "mpmc_pi_collective", "mpmc_debug_mark_moved", "mpmc_debug_pair_profile"]


KERNEL_CLASSES = ["energy_total", "pair", "structure", "field_recip", "field_real", "rank", "dipole_sweep", "gs_sweep", "palmo", "gs_precompute"]


def lib():
    """Load mpmcxx_b200/libmpmc_b200.so (building it first if the sources are newer)."""
    global _lib
    if _lib is None:
        path = _build.build_library()
        L = C.CDLL(path)
        vp = C.c_void_p
        L.mpmc_last_error.restype = C.c_char_p
        L.mpmc_create.argtypes = [C.POINTER(MpmcConfig), C.POINTER(vp)]
        L.mpmc_destroy.argtypes = [vp]
        L.mpmc_set_cell.argtypes = [vp, _dp]
        L.mpmc_get_cell.argtypes = [vp, _dp]
        L.mpmc_upload_sites.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
        L.mpmc_update_sites.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp]
        L.mpmc_update_sites_all_beads.argtypes = [vp, C.c_int, C.c_int, _dp]
        L.mpmc_insert_sites.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.mpmc_remove_sites.argtypes = [vp, C.c_int, C.c_int]
        L.mpmc_num_sites.argtypes = [vp, C.POINTER(C.c_int)]
        L.mpmc_energy.argtypes = [vp, C.POINTER(MpmcEnergyOut)]
        L.mpmc_energy_enqueue.argtypes = [vp]
        L.mpmc_energy_fetch.argtypes = [vp, C.POINTER(MpmcEnergyOut)]
        L.mpmc_download_dipoles.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.mpmc_download_rank_metric.argtypes = [vp, C.c_int, _dp]
        L.mpmc_pi_potential.argtypes = [vp, vp, _dp]
        L.mpmc_pi_chain.argtypes = [vp, C.c_int, C.POINTER(C.c_double), vp, vp, C.POINTER(C.c_int)]
        L.mpmc_nccl_get_unique_id.argtypes = [C.c_char_p]
        L.mpmc_nccl_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
        L.mpmc_pi_potential_allreduce.argtypes = [vp, C.c_int, _dp, C.POINTER(C.c_double)]
        L.mpmc_pi_chain_allreduce.argtypes = [vp, C.POINTER(C.c_double)]
        L.mpmc_debug_gs_profile.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
        L.mpmc_set_timing.argtypes = [vp, C.c_int]
        L.mpmc_get_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.mpmc_stream.argtypes = [vp]
        L.mpmc_stream.restype = vp
        L.mpmc_kernel_launches.argtypes = [vp]
        L.mpmc_kernel_launches.restype = C.c_longlong
        L.mpmc_probe_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mpmc_device_count.argtypes = [C.POINTER(C.c_int)]
        L.mpmc_pi_collective.argtypes = [vp]
        L.mpmc_debug_radial_table.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.c_int, _dp, C.c_void_p]
        L.mpmc_debug_cutoff_thresholds.argtypes = [C.c_double, _dp]
        L.mpmc_debug_mark_moved.argtypes = [vp, C.c_int, C.c_int]
        L.mpmc_debug_pair_profile.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _ck(rc):
    if rc:
        raise MpmcError(rc, lib().mpmc_last_error().decode())


def make_config(basis, opts: _config.EnergyOptions, n_beads=1, device=0, capacity=0) -> MpmcConfig:
    cfg = MpmcConfig()
    b = np.asarray(basis, dtype=np.float64).reshape(-1)
    for i in range(9):
        cfg.basis[i] = b[i]
    cfg.n_beads, cfg.device, cfg.capacity = n_beads, device, capacity
    for k in ("rd_lrc", "rd_only", "ewald_kmax", "ewald_alpha", "polarization", "polar_ewald", "polar_iterative", "damp_type", "polar_gs",
              "polar_gs_ranked", "polar_palmo", "polar_sor", "polar_esor", "polar_zodid", "polar_rrms", "polar_max_iter", "polar_damp",
              "polar_gamma", "polar_precision", "polar_ewald_alpha"):
        setattr(cfg, k, getattr(opts, k))
    return cfg


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _ck(lib().mpmc_nccl_get_unique_id(buf))
    return buf.raw


def radial_table(kind, param, u_lo, u_hi, u):
    """Host-side evaluation of the kernels' r^2-indexed table (kind 0: erfc(a r)/r; kind 1: the two real_term factors)."""
    u = np.ascontiguousarray(u, np.float64)
    o0, o1 = np.zeros_like(u), np.zeros_like(u)
    _ck(lib().mpmc_debug_radial_table(kind, param, u_lo, u_hi, u, u.size, o0, o1.ctypes.data_as(C.c_void_p)))
    return (o0, o1) if kind == 1 else o0


def cutoff_thresholds(cutoff):
    o = np.zeros(2)
    _ck(lib().mpmc_debug_cutoff_thresholds(cutoff, o))
    return o


def probe_fp64_peak(device=0):
    t, c = C.c_double(), C.c_double()
    _ck(lib().mpmc_probe_fp64_peak(device, C.byref(t), C.byref(c)))
    return t.value, c.value


class Engine:
    """One device-resident System (n_beads = 1) or a set of path-integral bead systems sharing a topology."""

    def __init__(self, system, beads=None, device=0, capacity=0, opts=None):
        self.system = system
        self.opts = opts if opts is not None else _config.from_keywords(system.opts)
        self.B = 1 if beads is None else int(beads.shape[0])
        cfg = make_config(system.basis, self.opts, self.B, device, capacity)
        h = C.c_void_p()
        _ck(lib().mpmc_create(C.byref(cfg), C.byref(h)))
        self.h = h
        pos = system.pos if beads is None else beads
        self.upload(system, pos)

    def close(self):
        if getattr(self, "h", None):
            lib().mpmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n(self):
        v = C.c_int()
        _ck(lib().mpmc_num_sites(self.h, C.byref(v)))
        return v.value

    def upload(self, system, pos):
        p = np.ascontiguousarray(pos, np.float64).reshape(-1)
        assert p.size == self.B * system.n * 3
        _ck(lib().mpmc_upload_sites(self.h, system.n, p, np.ascontiguousarray(system.charge), system.alpha, system.eps, system.sigma,
                                    system.mass, system.mol, system.frozen))

    def set_cell(self, basis):
        _ck(lib().mpmc_set_cell(self.h, np.ascontiguousarray(basis, np.float64).reshape(-1)))

    def cell(self):
        o = np.zeros(22)
        _ck(lib().mpmc_get_cell(self.h, o))
        return dict(basis=o[:9].reshape(3, 3).copy(), recip=o[9:18].reshape(3, 3).copy(), volume=o[18], cutoff=o[19],
                    ewald_alpha=o[20], polar_ewald_alpha=o[21])

    def update_sites(self, first, pos, bead=0):
        p = np.ascontiguousarray(pos, np.float64).reshape(-1)
        _ck(lib().mpmc_update_sites(self.h, bead, first, p.size // 3, p))

    def update_sites_all_beads(self, first, pos):
        p = np.ascontiguousarray(pos, np.float64)
        _ck(lib().mpmc_update_sites_all_beads(self.h, first, p.shape[1], p.reshape(-1)))

    def insert_sites(self, before, pos, charge, alpha, eps, sigma, mass, frozen=0):
        p = np.ascontiguousarray(pos, np.float64)
        count = p.size // (3 * self.B)
        f = lambda a: np.ascontiguousarray(a, np.float64)
        _ck(lib().mpmc_insert_sites(self.h, before, count, p.reshape(-1), f(charge), f(alpha), f(eps), f(sigma), f(mass), frozen))

    def remove_sites(self, first, count):
        _ck(lib().mpmc_remove_sites(self.h, first, count))

    def energy_all(self):
        out = (MpmcEnergyOut * self.B)()
        _ck(lib().mpmc_energy(self.h, out))
        return [o.as_dict() for o in out]

    def energy(self):
        return self.energy_all()[0]

    def enqueue(self):
        _ck(lib().mpmc_energy_enqueue(self.h))

    def fetch(self):
        out = (MpmcEnergyOut * self.B)()
        _ck(lib().mpmc_energy_fetch(self.h, out))
        return [o.as_dict() for o in out]

    def dipoles(self, bead=0):
        n = self.n
        arrs = [np.zeros((n, 3)) for _ in range(4)]
        _ck(lib().mpmc_download_dipoles(self.h, bead, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))
        rk = np.zeros(n)
        _ck(lib().mpmc_download_rank_metric(self.h, bead, rk))
        return dict(mu=arrs[0], ef_static=arrs[1], ef_induced=arrs[2], ef_induced_change=arrs[3], rank_metric=rk)

    def pi_potential(self):
        per = np.zeros((self.B, 4))
        sums = np.zeros(4)
        _ck(lib().mpmc_pi_potential(self.h, per.ctypes.data_as(C.c_void_p), sums))
        return per, sums

    def nccl_init(self, uid: bytes, rank: int, nranks: int):
        _ck(lib().mpmc_nccl_init(self.h, uid, rank, nranks))

    def pi_collective(self):
        return {0: "none (one GPU)", 1: "ncclAllReduce of 4 doubles", 2: "peer-memory mailboxes fused into the assembly kernel (k_pi_sums_xchg)"}[lib().mpmc_pi_collective(self.h)]

    def pi_potential_allreduce(self, P_global):
        means = np.zeros(4)
        pot = C.c_double()
        _ck(lib().mpmc_pi_potential_allreduce(self.h, P_global, means, C.byref(pot)))
        return pot.value, means

    def pi_chain_allreduce(self):
        v = C.c_double()
        _ck(lib().mpmc_pi_chain_allreduce(self.h, C.byref(v)))
        return v.value

    def pi_chain(self, closed=True):
        nmol = int(self.system.mol.max()) + 1 if self.n == self.system.n else None
        v, nm = C.c_double(), C.c_int()
        # molecule count is not known before the call when sites were inserted/removed: size generously
        cap = self.n
        com = np.zeros((self.B, cap, 3))
        mm = np.zeros(cap)
        _ck(lib().mpmc_pi_chain(self.h, 1 if closed else 0, C.byref(v), com.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), C.byref(nm)))
        k = nm.value
        com = com.reshape(-1)[: self.B * k * 3].reshape(self.B, k, 3).copy()
        return v.value, com, mm[:k].copy()

    def mark_moved(self, first, count):
        _ck(lib().mpmc_debug_mark_moved(self.h, first, count))

    def set_timing(self, on=True):
        _ck(lib().mpmc_set_timing(self.h, 1 if on else 0))

    def timing(self):
        """-> {kernel class: (total ms, launches)} accumulated since set_timing(True)"""
        ms = (C.c_double * len(KERNEL_CLASSES))()
        cnt = (C.c_longlong * len(KERNEL_CLASSES))()
        _ck(lib().mpmc_get_timing(self.h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_CLASSES)}

    def stream(self):
        return lib().mpmc_stream(self.h)

    def launches(self):
        return lib().mpmc_kernel_launches(self.h)
