"""Named parity cases shared by the golden generator, the oracle tests and the GPU parity tests."""
import json
import os

import numpy as np

from mpmcxx_b200 import workloads as W

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _with(s, **kw):
    s.opts.update({k: str(v) for k, v in kw.items()})
    return s


def _without(s, *keys):
    for k in keys:
        s.opts.pop(k, None)
    return s


CLASSIC = {
    # LJ only (SURVEY §8c KATs): lattice pairs sit exactly on the cutoff -> exercises the `rimg - 1e-12 < rc` test
    "lj_lattice_4": lambda: W.lj_lattice(4, 20.0),
    "lj_lattice_8": lambda: W.lj_lattice(8, 30.0),
    "lj_jitter_6_nolrc": lambda: _with(W.lj_lattice(6, 24.0, jitter=0.3, round4=False), rd_lrc="off", rd_only="on"),
    # 104-site polarizable + Ewald KAT, three solver variants
    "kat_gs_ranked_palmo": lambda: W.polar_kat(W.SOLVER_GS_RANKED_PALMO),
    "kat_jacobi10": lambda: W.polar_kat(W.SOLVER_JACOBI10),
    "kat_gs_precision": lambda: W.polar_kat(W.SOLVER_GS_PRECISION),
    # non-orthorhombic cell, unwrapped coordinates, null-parameter sites, mobile + frozen
    "tri_jacobi10": lambda: W.triclinic_mix(),
    "tri_gs_ranked_palmo": lambda: W.triclinic_mix(solver=W.SOLVER_GS_RANKED_PALMO),
    "tri_gs6": lambda: W.triclinic_mix(solver={"polar_gs": "on", "polar_max_iter": "6"}),
    "tri_sor": lambda: W.triclinic_mix(solver={"polar_sor": "on", "polar_gamma": "0.8", "polar_max_iter": "8"}),
    "tri_esor_rrms": lambda: W.triclinic_mix(solver={"polar_esor": "on", "polar_gamma": "1.2", "polar_max_iter": "8", "polar_rrms": "on"}),
    "tri_zodid": lambda: W.triclinic_mix(solver={"polar_zodid": "on"}),
    "tri_jacobi_precision": lambda: W.triclinic_mix(solver={"polar_precision": "1e-7"}),
    "tri_nopbc_field": lambda: _with(W.triclinic_mix(), polar_ewald="off"),
    "tri_no_polar": lambda: _with(_without(W.triclinic_mix(), "polar_max_iter"), polarization="off"),
    "tri_alpha_set": lambda: _with(W.triclinic_mix(), ewald_alpha="0.31", polar_ewald_alpha="0.27", ewald_kmax="5"),
    # scaled-down config 4: frozen framework + randomly oriented five-site H2
    "h2fw_6_gs_ranked_palmo": lambda: W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt"),
    "h2fw_6_jacobi10": lambda: W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_JACOBI10, ensemble="nvt"),
}

PI = {
    "pi_h2_single_27x8": lambda: W.pi_h2_cluster(n_side=3, P=8, L=40.0),
    "pi_h2_five_8x4": lambda: W.pi_h2_cluster(n_side=2, P=4, L=30.0, five_site=True),
}


def displaced(s, seed=99):
    """Move the last mobile molecule rigidly (a displace move, System.MonteCarlo.cpp:875) — deterministic."""
    rs = np.random.RandomState(seed)
    t = s.copy()
    m = int(t.mol[np.nonzero(t.frozen == 0)[0][-1]])
    t.pos[t.mol == m] += rs.uniform(-0.4, 0.4, size=3)
    return t


def load_golden(name):
    """-> (SiteSystem, dict of reference outputs) from tests/golden/<name>.npz"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    s = W.SiteSystem(z["basis"], z["pos"], z["charge_e"], z["alpha"], z["eps"], z["sigma"], z["mass"], z["mol"], z["frozen"],
                     [str(a) for a in z["atomtype"]], [str(a) for a in z["moltype"]], json.loads(str(z["opts"])))
    refd = {k: z[k] for k in z.files if k.startswith(("ref_", "cell_")) or k in ("moved_pos", "beads")}
    return s, refd
