"""Synthetic systems for parity tests and bench.py, plus writers for the reference's own input formats.

Everything here is host-side plumbing (numpy only).  A `SiteSystem` is the flat, list-ordered view
of the reference's Molecule -> Atom linked lists (src/System.cpp:881-904 `rebuild_arrays`): one row
per site, in the order `pairs()` walks them, which is also the order the engine's SoA uses.

The PQR / input-file writers follow the formats the reference parses
(src/System.cpp:515-770 `read_molecules`, src/SimulationControl.cpp:258-1616 `process_command`) so
that the very same system can be fed to the reference harness (oracle/_ref) and to the engine.

Configs are the ones SURVEY.md §8(d) / BASELINE.json name:
  lj_lattice / polar_kat        small known-answer systems
  lj_argon                      config 3: bulk LJ argon NVT, N=4096
  h2_framework                  config 4: frozen framework + five-site polarizable H2, N=10 000
  pi_h2_cluster                 config 5: path-integral H2 cluster, 512 molecules x 64 beads
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

E2REDUCED = 408.7816  # src/constants.h:32, applied to PQR charges at src/System.cpp:624


@dataclass
class SiteSystem:
    basis: np.ndarray                 # (3,3) rows = lattice vectors (basis1..3)
    pos: np.ndarray                   # (n,3) Angstrom, unwrapped
    charge_e: np.ndarray              # (n,) in e, as written in a PQR file
    alpha: np.ndarray                 # (n,) polarizability, A^3
    eps: np.ndarray                   # (n,) LJ epsilon, K
    sigma: np.ndarray                 # (n,) LJ sigma, A
    mass: np.ndarray                  # (n,) amu
    mol: np.ndarray                   # (n,) int32 molecule index, non-decreasing
    frozen: np.ndarray                # (n,) int32
    atomtype: list = field(default_factory=list)
    moltype: list = field(default_factory=list)
    opts: dict = field(default_factory=dict)   # input-file keywords (strings)

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    @property
    def charge(self) -> np.ndarray:
        """Charges in the reference's reduced units sqrt(K*A) (src/System.cpp:624)."""
        return self.charge_e * E2REDUCED

    def copy(self) -> "SiteSystem":
        return SiteSystem(self.basis.copy(), self.pos.copy(), self.charge_e.copy(), self.alpha.copy(),
                          self.eps.copy(), self.sigma.copy(), self.mass.copy(), self.mol.copy(),
                          self.frozen.copy(), list(self.atomtype), list(self.moltype), dict(self.opts))


def _mk(basis, pos, q, alpha, eps, sigma, mass, mol, frozen, atomtype, moltype, opts) -> SiteSystem:
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return SiteSystem(f(basis).reshape(3, 3), f(pos).reshape(-1, 3), f(q), f(alpha), f(eps), f(sigma), f(mass),
                      np.ascontiguousarray(np.asarray(mol, dtype=np.int32)),
                      np.ascontiguousarray(np.asarray(frozen, dtype=np.int32)), list(atomtype), list(moltype),
                      dict(opts))


# ----------------------------------------------------------------------------------------------
# writers for the reference's formats
# ----------------------------------------------------------------------------------------------
def write_pqr(system: SiteSystem, path: str) -> None:
    """PQR as src/System.cpp:583-587 scans it: ATOM id atomtype moltype F|M molid x y z mass q alpha eps sigma
    omega gwp_alpha c6 c8 c10 c9.  Doubles are written %.17g so the reference parses the exact bits."""
    with open(path, "w") as fp:
        for i in range(system.n):
            fz = "F" if system.frozen[i] else "M"
            x, y, z = system.pos[i]
            fp.write("ATOM %d %s %s %s %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g 0.0 0.0 0.0 0.0 0.0 0.0\n" % (
                i + 1, system.atomtype[i], system.moltype[i], fz, int(system.mol[i]) + 1, x, y, z,
                system.mass[i], system.charge_e[i], system.alpha[i], system.eps[i], system.sigma[i]))
        fp.write("END\n")


DEFAULT_OPTS = {
    "job_name": "job", "ensemble": "nvt", "temperature": "77.0", "numsteps": "10", "corrtime": "10",
    "seed": "1", "move_factor": "0.01", "rot_factor": "0.01", "wrapall": "on", "pop_histogram": "off",
    "traj_output": "off", "parallel_restarts": "off", "feynman_hibbs": "off", "h2_fugacity": "off",
}


def write_input(system: SiteSystem, path: str, pqr_name: str) -> None:
    opts = dict(DEFAULT_OPTS)
    opts.update(system.opts)
    with open(path, "w") as fp:
        for k, v in opts.items():
            if not k.startswith("_"):
                fp.write("%-30s %s\n" % (k, v))
        for line in opts.get("_lines", ()):      # keywords that may repeat (sorbate_bondlength <type> <A>, ...), in the order given
            fp.write(line + "\n")
        for i in range(3):
            fp.write("basis%d %.17g %.17g %.17g\n" % (i + 1, *system.basis[i]))
        fp.write("pqr_input %s\n" % pqr_name)


def write_reference_job(system: SiteSystem, directory: str) -> str:
    """Write input.in + input.pqr into `directory`; returns the input-file path."""
    os.makedirs(directory, exist_ok=True)
    write_pqr(system, os.path.join(directory, "input.pqr"))
    inp = os.path.join(directory, "input.in")
    write_input(system, inp, "input.pqr")
    return inp


# ----------------------------------------------------------------------------------------------
# small known-answer systems
# ----------------------------------------------------------------------------------------------
AR = dict(eps=119.8, sigma=3.405, mass=39.948)


def lj_lattice(g: int, L: float, jitter: float = 0.0, seed: int = 12345, round4: bool = True) -> SiteSystem:
    """Simple-cubic g^3 argon sites at ((i+1/2)L/g - L/2), rd_lrc on (SURVEY §8c 'LJ lattice' KAT)."""
    idx = np.arange(g)
    c = (idx + 0.5) * L / g - L / 2
    pos = np.stack(np.meshgrid(c, c, c, indexing="ij"), axis=-1).reshape(-1, 3)
    if jitter:
        rs = np.random.RandomState(seed)
        pos = pos + (rs.random_sample(pos.shape) * 2.0 - 1.0) * jitter
    if round4:
        pos = np.round(pos, 4)
    n = g ** 3
    opts = {"ensemble": "nvt", "temperature": "87.0", "polarization": "off", "rd_lrc": "on"}
    return _mk(np.eye(3) * L, pos, np.zeros(n), np.zeros(n), np.full(n, AR["eps"]), np.full(n, AR["sigma"]),
               np.full(n, AR["mass"]), np.arange(n), np.zeros(n), ["Ar"] * n, ["Ar"] * n, opts)


def lj_argon(n_side: int = 16, L: float = 60.0) -> SiteSystem:
    """Config 3: bulk LJ argon NVT (N = n_side^3 = 4096, rho* = 0.75), lattice + uniform +-0.30 A jitter."""
    s = lj_lattice(n_side, L, jitter=0.30, seed=12345, round4=False)
    s.opts.update({"move_factor": "0.01", "seed": "7", "rd_only": "on"})
    return s


# five-site polarizable H2 (BSSP-like), SURVEY §8c: name, offset along axis, q(e), alpha, eps, sigma, mass
H2_SITES = [
    ("H2G", 0.000, -0.7464, 0.6938, 12.76532, 3.15528, 0.0),
    ("H2E", 0.371, 0.3732, 0.00044, 0.0, 0.0, 1.008),
    ("H2E", -0.371, 0.3732, 0.00044, 0.0, 0.0, 1.008),
    ("H2N", 0.329, 0.0, 0.0, 2.16726, 2.37031, 0.0),
    ("H2N", -0.329, 0.0, 0.0, 2.16726, 2.37031, 0.0),
]

POLAR_OPTS = {
    "polarization": "on", "polar_damp_type": "exponential", "polar_damp": "2.1304", "polar_ewald": "on",
    "polar_iterative": "on", "ewald_kmax": "7", "rd_lrc": "on",
}
SOLVER_GS_RANKED_PALMO = {"polar_gs_ranked": "on", "polar_palmo": "on", "polar_gamma": "1.03", "polar_max_iter": "4"}
SOLVER_JACOBI10 = {"polar_max_iter": "10"}
SOLVER_GS_PRECISION = {"polar_gs": "on", "polar_precision": "1e-9"}


def _random_axes(rs: np.random.RandomState, m: int) -> np.ndarray:
    v = rs.normal(size=(m, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def h2_framework(ncell: int = 20, a: float = 4.0, n_h2: int = 400, seed: int = 2024, solver: dict | None = None,
                 random_axes: bool = True, ensemble: str = "uvt") -> SiteSystem:
    """Config 4: ncell^3 frozen framework sites (simple cubic, spacing a, q = +-0.5 e alternating,
    alpha 1.2886, eps 30, sigma 3.4, ONE frozen molecule) + n_h2 five-site H2 at body centres of distinct cells.
    ncell=20, n_h2=400 -> N = 10 000."""
    L = ncell * a
    rs = np.random.RandomState(seed)
    ii = np.arange(ncell)
    I, J, K = np.meshgrid(ii, ii, ii, indexing="ij")
    fpos = np.stack([(I + 0.25) * a - L / 2, (J + 0.25) * a - L / 2, (K + 0.25) * a - L / 2], axis=-1).reshape(-1, 3)
    par = ((I + J + K) % 2).reshape(-1)
    nf = ncell ** 3
    q = list(np.where(par == 0, 0.5, -0.5))
    pos = [fpos]
    alpha = [1.2886] * nf
    eps = [30.0] * nf
    sig = [3.4] * nf
    mass = [12.011] * nf
    mol = [0] * nf
    frozen = [1] * nf
    at = ["C"] * nf
    mt = ["MOF"] * nf
    cells = rs.permutation(nf)[:n_h2]
    centres = np.stack(np.unravel_index(cells, (ncell, ncell, ncell)), axis=-1).astype(np.float64)
    centres = (centres + 0.75) * a - L / 2
    axes = _random_axes(rs, n_h2) if random_axes else np.tile(np.array([1.0, 0.0, 0.0]), (n_h2, 1))
    for m in range(n_h2):
        for (name, off, qq, al, ee, ss, ms) in H2_SITES:
            pos.append((centres[m] + off * axes[m])[None, :])
            q.append(qq); alpha.append(al); eps.append(ee); sig.append(ss); mass.append(ms)
            mol.append(1 + m); frozen.append(0); at.append(name); mt.append("H2")
    opts = dict(POLAR_OPTS)
    opts.update({"ensemble": ensemble, "temperature": "77.0", "move_factor": "0.05", "rot_factor": "0.05"})
    if ensemble == "uvt":
        opts.update({"pressure": "1.0", "h2_fugacity": "on", "insert_probability": "0.3", "free_volume": "%.1f" % (L ** 3)})
    opts.update(solver if solver is not None else SOLVER_GS_RANKED_PALMO)
    return _mk(np.eye(3) * L, np.concatenate(pos, axis=0), q, alpha, eps, sig, mass, mol, frozen, at, mt, opts)


def uvt_pore(ncell: int = 5, n_h2: int = 6, seed: int = 3, pressure: float = 5.0, solver: dict | None = None) -> SiteSystem:
    """Small grand-canonical test system: the config-4 framework with an open channel along z (so that insertions can be
    accepted) and a few five-site H2 inside it; uVT with fugacity = pressure (no equation of state)."""
    s = h2_framework(ncell=ncell, n_h2=n_h2, solver=solver if solver is not None else SOLVER_GS_RANKED_PALMO, ensemble="uvt", seed=seed)
    L = ncell * 4.0
    keep = ~((s.frozen == 1) & (np.abs(s.pos[:, 0]) < 5.5) & (np.abs(s.pos[:, 1]) < 5.5))
    rs = np.random.RandomState(seed)
    pos = s.pos.copy()
    for m in range(1, n_h2 + 1):
        idx = np.nonzero(s.mol == m)[0]
        newc = np.array([rs.uniform(-3, 3), rs.uniform(-3, 3), -L / 2 + (m - 0.5) * L / n_h2])
        pos[idx] += newc - pos[idx[0]]
    t = SiteSystem(s.basis, pos[keep], s.charge_e[keep], s.alpha[keep], s.eps[keep], s.sigma[keep], s.mass[keep], s.mol[keep], s.frozen[keep],
                   [a for a, k in zip(s.atomtype, keep) if k], [a for a, k in zip(s.moltype, keep) if k], dict(s.opts))
    t.opts.update({"pressure": str(pressure), "h2_fugacity": "off", "insert_probability": "0.4", "move_factor": "0.05", "rot_factor": "0.05"})
    return t


def argon_dimer_pi() -> SiteSystem:
    """sample-input/pi001-argon-dimer-2K as shipped by the reference (Ar-Ar-4A.pqr + equilibrate.in), rebuilt here: two LJ argon
    atoms 4 A apart in a 10^4 A box at 2 K, bead perturbation probability 0.9, trial chain length 4."""
    opts = {"job_name": "ArAr2K", "ensemble": "pi_nvt", "temperature": "2.0", "polarization": "off", "numsteps": "100000", "corrtime": "25",
            "seed": "1", "move_factor": "0.03", "rot_factor": "1.0", "bead_perturb_probability": "0.90", "free_volume": "1.0e12",
            "PI_trial_chain_length": "4", "pqr_restart": "ArAr2K.restart.pqr"}
    pos = np.array([[-2.0, 0.0, 0.0], [2.0, 0.0, 0.0]])
    return _mk(np.eye(3) * 10000.0, pos, np.zeros(2), np.zeros(2), np.full(2, AR["eps"]), np.full(2, AR["sigma"]), np.full(2, AR["mass"]),
               [0, 1], [0, 0], ["Ar", "Ar"], ["Ar", "Ar"], opts)


def polar_kat(solver: dict | None = None) -> SiteSystem:
    """104-site polarizable + Ewald known-answer system of SURVEY §8c: 4^3 frozen sites + 8 H2 on axis x, L = 16."""
    ncell, a, L = 4, 4.0, 16.0
    s = h2_framework(ncell=ncell, a=a, n_h2=0, solver=solver, ensemble="nvt")
    pos = [s.pos]; q = list(s.charge_e); alpha = list(s.alpha); eps = list(s.eps); sig = list(s.sigma)
    mass = list(s.mass); mol = list(s.mol); frozen = list(s.frozen); at = list(s.atomtype); mt = list(s.moltype)
    m = 0
    for i in (0, 2):
        for j in (0, 2):
            for k in (0, 2):
                c = np.array([(i + 0.75) * a - L / 2, (j + 0.75) * a - L / 2, (k + 0.75) * a - L / 2])
                for (name, off, qq, al, ee, ss, ms) in H2_SITES:
                    pos.append((c + np.array([off, 0.0, 0.0]))[None, :])
                    q.append(qq); alpha.append(al); eps.append(ee); sig.append(ss); mass.append(ms)
                    mol.append(1 + m); frozen.append(0); at.append(name); mt.append("H2")
                m += 1
    return _mk(s.basis, np.concatenate(pos, axis=0), q, alpha, eps, sig, mass, mol, frozen, at, mt, s.opts)


def triclinic_mix(n_mol: int = 24, seed: int = 5, solver: dict | None = None) -> SiteSystem:
    """Small non-orthorhombic cell with mobile three-site charged/polarizable molecules and a few frozen sites:
    exercises the general min-image path, rint ties, exclusions and null-parameter sites."""
    rs = np.random.RandomState(seed)
    basis = np.array([[14.0, 0.0, 0.0], [3.0, 13.0, 0.0], [-2.0, 2.5, 15.0]])
    pos = []; q = []; alpha = []; eps = []; sig = []; mass = []; mol = []; frozen = []; at = []; mt = []
    nfz = 12
    for i in range(nfz):
        f = np.array([(i % 3 + 0.5) / 3.0, ((i // 3) % 2 + 0.5) / 2.0, (i // 6 + 0.5) / 2.0])
        pos.append(f @ basis); q.append(0.4 if i % 2 == 0 else -0.4); alpha.append(1.1); eps.append(25.0)
        sig.append(3.2); mass.append(12.0); mol.append(0); frozen.append(1); at.append("C"); mt.append("FRM")
    inv = np.linalg.inv(basis)

    def too_close(p):
        if not pos:
            return False
        d = (np.array(pos) - p) @ inv
        d -= np.rint(d)
        return bool((np.linalg.norm(d @ basis, axis=1) < 2.6).any())

    for m in range(n_mol):
        while True:
            c = (rs.random_sample(3) * 3.0 - 1.0) @ basis     # deliberately spills outside the cell: unwrapped coords
            ax = _random_axes(rs, 1)[0]
            if not any(too_close(c + off * ax) for off in (0.0, 0.9, -0.9)):
                break
        for (name, off, qq, al, ee, ss, ms) in (("OA", 0.0, -0.6, 0.85, 60.0, 3.0, 16.0), ("HA", 0.9, 0.3, 0.0, 0.0, 0.0, 1.0),
                                               ("HB", -0.9, 0.3, 0.3, 8.0, 2.2, 1.0)):
            pos.append(c + off * ax); q.append(qq); alpha.append(al); eps.append(ee); sig.append(ss); mass.append(ms)
            mol.append(1 + m); frozen.append(0); at.append(name); mt.append("W")
    opts = dict(POLAR_OPTS)
    opts.update({"ensemble": "nvt", "temperature": "150.0"})
    opts.update(solver if solver is not None else SOLVER_JACOBI10)
    return _mk(basis, np.array(pos), q, alpha, eps, sig, mass, mol, frozen, at, mt, opts)


# ----------------------------------------------------------------------------------------------
# path-integral cluster (config 5)
# ----------------------------------------------------------------------------------------------
H2_SINGLE = dict(eps=34.2, sigma=2.96, mass=2.016)


def pi_h2_cluster(n_side: int = 8, P: int = 64, a: float = 3.8, L: float = 100.0, five_site: bool = False,
                  seed: int = 1, bead_sigma: float = 0.25, polarizable: bool = False, solver: dict | None = None):
    """Config 5: n_side^3 H2 molecules on a centred cubic lattice, P beads each.  Returns (template, beads) where
    `template` is the SiteSystem of bead 0's centroid geometry and `beads` is a (P, n, 3) array of per-bead site
    positions (centroid + one gaussian COM offset per molecule per bead; rigid molecules keep their orientation)."""
    rs = np.random.RandomState(seed)
    idx = np.arange(n_side)
    c = (idx - (n_side - 1) / 2.0) * a
    centres = np.stack(np.meshgrid(c, c, c, indexing="ij"), axis=-1).reshape(-1, 3)
    M = centres.shape[0]
    opts = {"ensemble": "pi_nvt", "temperature": "20.0", "polarization": "off", "rd_lrc": "on", "move_factor": "0.01",
            "bead_perturb_probability": "0.5", "PI_trial_chain_length": str(max(1, P // 4)), "free_volume": "1.0e6",
            "pqr_restart": "job.restart.pqr"}
    if not five_site:
        opts["rd_only"] = "on"
        tmpl = _mk(np.eye(3) * L, centres, np.zeros(M), np.zeros(M), np.full(M, H2_SINGLE["eps"]),
                   np.full(M, H2_SINGLE["sigma"]), np.full(M, H2_SINGLE["mass"]), np.arange(M), np.zeros(M),
                   ["H2"] * M, ["H2"] * M, opts)
    else:
        axes = _random_axes(rs, M)
        pos = []; q = []; alpha = []; eps = []; sig = []; mass = []; mol = []; at = []
        for m in range(M):
            for (name, off, qq, al, ee, ss, ms) in H2_SITES:
                pos.append(centres[m] + off * axes[m]); q.append(qq); alpha.append(al if polarizable else 0.0); eps.append(ee); sig.append(ss)
                mass.append(ms); mol.append(m); at.append(name)
        n = len(q)
        opts["ewald_kmax"] = "7"
        if polarizable:     # every bead system is polarized on its own (src/SimulationControl.PathIntegral.cpp:770-780)
            opts.update(POLAR_OPTS)
            opts.update(solver if solver is not None else SOLVER_GS_RANKED_PALMO)
        tmpl = _mk(np.eye(3) * L, np.array(pos), q, alpha, eps, sig, mass, mol, np.zeros(n), at, ["H2"] * n, opts)
    off = rs.normal(scale=bead_sigma, size=(P, M, 3))
    beads = tmpl.pos[None, :, :] + off[:, tmpl.mol, :]
    return tmpl, np.ascontiguousarray(beads)
