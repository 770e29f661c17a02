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
    # the other two Thole damping forms of thole_amatrix (src/System.Energy.cpp:2714-2731): linear and none
    "tri_linear_jacobi10": lambda: _with(W.triclinic_mix(), polar_damp_type="linear", polar_damp="1.662"),
    "tri_linear_gs_ranked_palmo": lambda: _with(W.triclinic_mix(solver=W.SOLVER_GS_RANKED_PALMO), polar_damp_type="linear", polar_damp="1.662"),
    "tri_damp_off_gs6": lambda: _with(W.triclinic_mix(solver={"polar_gs": "on", "polar_max_iter": "6"}), polar_damp_type="off"),
    "h2fw_6_linear_gs_ranked": lambda: _with(W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt"), polar_damp_type="linear", polar_damp="1.662"),
    # scaled-down config 4: frozen framework + randomly oriented five-site H2
    "h2fw_6_gs_ranked_palmo": lambda: W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt"),
    "h2fw_6_jacobi10": lambda: W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_JACOBI10, ensemble="nvt"),
    # 879 polarizable sites = 13.7 Gauss-Seidel blocks of 64: beyond the cluster's lookahead, so the updaters' flags, the far rows and
    # the wrapped (already swept) rows all take part; not a multiple of 64
    "h2fw_9_gs_ranked_palmo": lambda: W.h2_framework(ncell=9, n_h2=50, solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt"),
    # the 128-iteration fail-out of thole_iterative (src/System.Energy.cpp:3483-3494): over-relaxed SOR diverges, precision mode
    # never converges -> mu = alpha E_static, ef_induced_change = 0, iterator_failed = 1
    "tri_sor_failed": lambda: W.triclinic_mix(solver={"polar_sor": "on", "polar_gamma": "2.2", "polar_precision": "1e-6"}),
}

PI = {
    "pi_h2_single_27x8": lambda: W.pi_h2_cluster(n_side=3, P=8, L=40.0),
    "pi_h2_five_8x4": lambda: W.pi_h2_cluster(n_side=2, P=4, L=30.0, five_site=True),
    # polarizable bead systems: Gauss-Seidel ranked + Palmo per bead (81 polarizable sites = 2 blocks), and a precision-controlled
    # Jacobi solve whose beads converge after different numbers of iterations
    "pi_h2_polar_gs_27x4": lambda: W.pi_h2_cluster(n_side=3, P=4, L=30.0, five_site=True, polarizable=True, solver=W.SOLVER_GS_RANKED_PALMO),
    "pi_h2_polar_prec_8x4": lambda: W.pi_h2_cluster(n_side=2, P=4, L=30.0, five_site=True, polarizable=True, solver={"polar_precision": "1e-8"}, bead_sigma=0.5),
}


# seeded Markov chains whose accept/reject trajectory must reproduce the reference's: name -> (builder, P, steps)
def _traj_lj():
    s = W.lj_lattice(6, 24.0, jitter=0.3, round4=False)
    s.opts.update({"ensemble": "nvt", "temperature": "87.0", "seed": "7", "move_factor": "0.02", "rot_factor": "0.1", "numsteps": "10000"})
    return s


def _traj_kat():
    s = W.polar_kat(W.SOLVER_GS_RANKED_PALMO)
    s.opts.update({"ensemble": "nvt", "seed": "5", "move_factor": "0.06", "rot_factor": "0.1", "numsteps": "10000"})
    return s


def _traj_uvt():
    s = W.uvt_pore()
    s.opts.update({"seed": "11", "numsteps": "10000"})
    return s


def _traj_pi_h2():
    tmpl, _ = W.pi_h2_cluster(n_side=3, P=8, L=40.0)
    tmpl.opts.update({"seed": "3", "numsteps": "10000", "PI_trial_chain_length": "3"})
    return tmpl


def _traj_pi_orient(dummy_first):
    """Five-site H2, 8 molecules x 8 beads, with the orientational degree of freedom configured (sorbate_orientation_site /
    sorbate_bondlength / sorbate_reducedMass, src/SimulationControl.cpp:306-339; bead orientations by recursive bisection,
    src/SimulationControl.PathIntegral.cpp:1559-1697; orientational term of the acceptance, :978-1039, :490-547).  The reference
    uses the INDEX of the type's metadata record as the handle site (src/SimulationControl.cpp:2996-3004): with H2 alone that is
    site 0 (the massless-free centre site ON the COM: no rotation, zero bond vectors); with another type declared first it is
    site 1 (an H2E site 0.371 A from the COM: the beads really turn)."""
    def build():
        tmpl, _ = W.pi_h2_cluster(n_side=2, P=8, L=30.0, five_site=True)
        lines = ["sorbate_bondlength XX 1.0"] if dummy_first else []
        lines += ["sorbate_orientation_site H2 1", "sorbate_bondlength H2 0.742", "sorbate_reducedMass H2 8.368618e-28"]
        tmpl.opts.update({"seed": "17", "numsteps": "10000", "PI_trial_chain_length": "3", "rot_factor": "5.0", "_lines": lines})
        return tmpl
    return build


TRAJ = {
    "traj_nvt_lj216": (_traj_lj, 0, 10000),
    "traj_nvt_kat_gs_ranked": (_traj_kat, 0, 10000),
    "traj_uvt_pore": (_traj_uvt, 0, 10000),
    "traj_pi_argon_dimer": (W.argon_dimer_pi, 8, 10000),
    "traj_pi_h2_27x8": (_traj_pi_h2, 8, 10000),
    "traj_pi_h2_orient_site1": (_traj_pi_orient(True), 8, 10000),
    "traj_pi_h2_orient_site0": (_traj_pi_orient(False), 8, 4000),
}


# sample directories shipped with the reference, run UNMODIFIED (file texts are stored in the golden next to the reference's trajectory):
# name -> (directory under sample-input/, files the job reads, input file, Trotter number, steps)
SHIPPED = {
    # BASELINE config 1: two free argon atoms (eps = 0) at 2 K, potential == 0 -> the kinetic-energy series of the bead-spring path
    # (src/SimulationControl.PathIntegral.cpp:810-828) and the host loop; PQR with CRYST1 / BOX pseudo-molecule / CONECT records
    "shipped_pi000_free_argon": ("pi000-free-argon-2K", ["equilibrate.in", "Ar.pqr"], "equilibrate.in", 8, 10000),
    # the same directory's production input (`parallel_restarts on`): every bead system restarts from its own shipped restart file
    "shipped_pi000_input_restarts": ("pi000-free-argon-2K", ["input.in"] + ["Ar2K.restart-%04d.pqr" % i for i in range(4)] + ["Ar2K.restart-0000.pqr.last"],
                                     "input.in", 8, 10000),
    # BASELINE config 2: the argon dimer (LJ pair + bead springs), as shipped
    "shipped_pi001_argon_dimer": ("pi001-argon-dimer-2K", ["equilibrate.in", "Ar-Ar-4A.pqr"], "equilibrate.in", 8, 10000),
}


# PQR files as the reference writes them (System::write_molecules, src/System.Output.cpp:900-1091): name -> (builder, P, system index)
WRITTEN = {
    "tri_gs_ranked_palmo": (lambda: W.triclinic_mix(solver=W.SOLVER_GS_RANKED_PALMO), 0, -1),     # triclinic cell, unwrapped input, PDB-style coordinates
    "h2fw_6_jacobi10": (lambda: W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_JACOBI10, ensemble="nvt"), 0, -1),
    "pi_argon_dimer": (W.argon_dimer_pi, 8, 3),                                                    # 10^4 A cell: extended coordinates, file names with -0003
    "lj_nowrap": (lambda: _with(W.lj_lattice(4, 20.0, jitter=3.0, round4=False), wrapall="off"), 0, -1),
}


def load_shipped(name):
    """-> ({file name: text}, input file name, P, reference trajectory [steps, 5])"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    files = {str(k): str(v) for k, v in zip(z["file_names"], z["file_texts"])}
    return files, str(z["input_name"]), int(z["P"]), z["traj"]


def write_shipped(files, directory):
    os.makedirs(directory, exist_ok=True)
    for k, v in files.items():
        with open(os.path.join(directory, k), "w") as fp:
            fp.write(v)


# ensemble averages (north_star's third criterion): name -> (builder, P, steps, seeds run by the reference for the golden, seeds the
# engine's chains use in the test).  Different seeds on the two sides: the chains are statistically independent, and the means of
# energy / N (uVT) and of the potential / kinetic energy (pi_nvt: kinetic = const - 1/2 omega^2 sum m <r^2> of the bead chains,
# src/SimulationControl.PathIntegral.cpp:810-828) must agree within the combined standard error.
def _avg_uvt():
    s = W.uvt_pore()
    s.opts.update({"numsteps": "6000"})
    return s


def _avg_pi():
    tmpl, _ = W.pi_h2_cluster(n_side=3, P=8, L=40.0)
    tmpl.opts.update({"numsteps": "6000", "PI_trial_chain_length": "3"})
    return tmpl


AVERAGES = {
    "avg_uvt_pore": (_avg_uvt, 0, 6000, (101, 102, 103, 104), (201, 202, 203, 204)),
    "avg_pi_h2_27x8": (_avg_pi, 8, 6000, (101, 102, 103, 104), (201, 202, 203, 204)),
}
AVG_BLOCKS = 6


# jobs whose parse (input file + PQR -> flat site table + cell) is pinned against the reference's own readers: (builder, Trotter number)
PARSED = {
    "lj_lattice_4": (lambda: W.lj_lattice(4, 20.0), 0),
    "tri_gs_ranked_palmo": (lambda: W.triclinic_mix(solver=W.SOLVER_GS_RANKED_PALMO), 0),
    "h2fw_6_jacobi10": (lambda: W.h2_framework(ncell=6, n_h2=20, solver=W.SOLVER_JACOBI10, ensemble="nvt"), 0),
    "tri_alpha_set": (lambda: _with(W.triclinic_mix(), ewald_alpha="0.31", polar_ewald_alpha="0.27", ewald_kmax="5"), 0),
    "pi_h2_five_8x4": (lambda: W.pi_h2_cluster(n_side=2, P=4, L=30.0, five_site=True)[0], 4),
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


def load_golden_traj(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    s = W.SiteSystem(z["basis"], z["pos"], z["charge_e"], z["alpha"], z["eps"], z["sigma"], z["mass"], z["mol"], z["frozen"],
                     [str(a) for a in z["atomtype"]], [str(a) for a in z["moltype"]], json.loads(str(z["opts"])))
    return s, {"P": z["P"], "traj": z["traj"]}


# malformed jobs and the error code the reference throws for them when it reads the job (main.cpp:54-63 catches the int): name ->
# (builder, P, mutation of the SiteSystem).  tests/golden/input_errors.json holds the reference's codes.
def _pi_small():
    return W.pi_h2_cluster(n_side=2, P=4, L=30.0, five_site=False)[0]


def _opt(**kw):
    def m(s):
        for k, v in kw.items():
            if v is None:
                s.opts.pop(k, None)
            else:
                s.opts[k] = v
    return m


def _line(text):
    def m(s):
        s.opts["_lines"] = list(s.opts.get("_lines", ())) + [text]
    return m


INPUT_ERRORS = {
    "ok_nvt": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="5")),
    "zero_temperature": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="5", temperature="0.0")),
    "zero_numsteps": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="0")),
    "zero_corrtime": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="5", corrtime="0")),
    "missing_argument": (lambda: W.lj_lattice(3, 20.0), 0, _line("temperature")),
    "unknown_keyword": (lambda: W.lj_lattice(3, 20.0), 0, _line("no_such_keyword on")),
    "retired_rot_probability": (lambda: W.lj_lattice(3, 20.0), 0, _line("rot_probability 0.5")),
    "retired_move_probability": (lambda: W.lj_lattice(3, 20.0), 0, _line("move_probability 0.5")),
    "bad_switch_value": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="5", rd_only="maybe")),
    "bad_ensemble": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="5", ensemble="nonsense")),
    "uvt_without_pressure": (lambda: W.lj_lattice(3, 20.0), 0, _opt(seed="1", numsteps="5", ensemble="uvt", insert_probability="0.3", free_volume="8000.0")),
    "pi_three_beads": (_pi_small, 3, _opt(seed="1", numsteps="5", PI_trial_chain_length="1")),
    "pi_two_beads": (_pi_small, 2, _opt(seed="1", numsteps="5", PI_trial_chain_length="1")),
    "pi_chain_too_long": (_pi_small, 4, _opt(seed="1", numsteps="5", PI_trial_chain_length="4")),
    "pi_chain_missing": (_pi_small, 4, _opt(seed="1", numsteps="5", PI_trial_chain_length=None)),
    "pi_ok": (_pi_small, 4, _opt(seed="1", numsteps="5", PI_trial_chain_length="2")),
}


# System::update_root_averages (src/System.Averages.cpp:8-208): the job and the synthetic samples behind tests/golden/root_averages.npz
def uvt_pore_for_averages():
    s = W.uvt_pore()
    s.opts.update({"h2_fugacity": "on"})        # a fugacity keyword fills fugacities[0], which the excess adsorption uses
    return s


def root_average_samples(n=300, seed=3):
    """[n, 6] = (energy, coulombic, rd, polarization, N, NU = N * energy): N wanders between 20 and 40, the energy follows it with noise"""
    import numpy as _np
    rs = _np.random.RandomState(seed)
    x = _np.zeros((n, 6))
    x[:, 4] = rs.randint(20, 40, size=n)
    x[:, 0] = -50.0 * x[:, 4] + rs.normal(scale=30.0, size=n)
    x[:, 1] = 0.3 * x[:, 0]; x[:, 2] = 0.6 * x[:, 0]; x[:, 3] = 0.1 * x[:, 0]
    x[:, 5] = x[:, 0] * x[:, 4]
    return x


# classic chains with the reference's averaging every correlation time (System::update_root_averages inside the mc loop):
# name -> (builder, steps, corrtime); tests/golden/mc_averages.npz holds what the unmodified reference accumulates
MC_AVERAGES = {
    "nvt_lj216": (_traj_lj, 2000, 10),
    "uvt_pore": (_traj_uvt, 1500, 10),
}

# path-integral chains with the reference's averaging (initial state + every correlation time + the end): name -> (builder, P, steps, corrtime)
PI_AVERAGES = {
    "pi_h2_27x8": (_traj_pi_h2, 8, 600, 10),
    "pi_argon_dimer": (W.argon_dimer_pi, 8, 1500, 25),
}

# the geometry a chain ends with, as the PQR file the reference writes for it (final / restart state after real moves, insertions and
# removals included): name -> (builder, P, steps, bead system written); texts in tests/golden/final_pqr.npz
FINAL_PQR = {
    "uvt_pore": (_traj_uvt, 0, 400, -1),
    "nvt_lj216": (_traj_lj, 0, 300, -1),
    "pi_h2_27x8": (_traj_pi_h2, 8, 200, 3),
}
