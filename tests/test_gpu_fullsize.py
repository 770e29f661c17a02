"""GPU parity at the sizes bench.py measures: every BASELINE configuration at its real size against the CPU oracle
(oracle/oracle.c, pinned to the reference in tests/test_oracle.py; ~8 s per evaluation at N = 10 000), the benchmarked solver
included — Gauss-Seidel ranked x 4 + Palmo over 144 blocks of 64 polarizable sites, where the pipeline's ring of 5 slots, its
lookahead of 4 blocks, the 7 helpers and the updater kernel all interact.

Bar (north_star): every energy sub-term within 1e-10 relative in FP64; iteration counts and the integer rank metric exact;
dipoles and fields within 1e-10 of the largest component.  The cancelled Ewald total is judged against the sum of |sub-terms|.
"""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-10
SUBTERMS = (("rd_pair", "rd"), ("rd_lrc_pair", "lrc_pair"), ("rd_lrc_self", "lrc_self"), ("es_real", "es_real"),
            ("es_self_intra", "es_self_intra"), ("es_reciprocal", "es_recip"), ("es_self", "es_self"), ("polarization_energy", "polar"))


def _eng():
    from mpmcxx_b200 import engine
    return engine


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def _compare_with_oracle(e, s, sites=True):
    from oracle import port
    o = e.energy()
    p = port.energy(s, want_sites=sites)
    # A pair sum whose terms cancel (config 4: es_real = 0.87 K out of 3.3e8 K of |q_i q_j erfc(a r)/r|) is only defined to the rounding of
    # its terms: every term carries a relative error of a few 2^-53 on both sides, and the reference's own sequential sum depends on its
    # list order at that level.  Tolerance per sub-term: 1e-10 of the value + ONE ulp (2^-52) of the sum of |terms|,
    # which the oracle reports next to each pair sum.
    noise = {"rd": p["rd_abs"], "es_real": p["es_real_abs"], "es_self_intra": p["es_intra_abs"]}
    for k_o, k_p in SUBTERMS:
        tol = RTOL * abs(p[k_p]) + 2.0 ** -52 * noise.get(k_p, 0.0)
        if p[k_p] == 0.0:
            assert abs(o[k_o]) < 1e-9, (k_o, o[k_o])
        else:
            assert abs(o[k_o] - p[k_p]) <= tol, (k_o, o[k_o], p[k_p], _rel(o[k_o], p[k_p]))
    scale = abs(o["es_real"]) + abs(o["es_self_intra"]) + abs(o["es_reciprocal"]) + abs(o["es_self"])
    assert abs(o["coulombic_energy"] - p["coulombic"]) <= RTOL * max(scale, 1e-300)
    assert abs(o["rd_energy"] - p["rd_total"]) <= RTOL * abs(p["rd_total"]) + 2.0 ** -52 * p["rd_abs"]
    assert o["n_pairs_in_cutoff"] == p["n_pairs_in_cutoff"]
    assert o["polarization_iterations"] == int(p["iterations"]) and o["iterator_failed"] == int(p["iterator_failed"])
    if sites and s.opts.get("polarization") == "on":
        d = e.dipoles()
        for k in ("mu", "ef_static", "ef_induced"):
            err = np.abs(d[k] - p[k]).max() / np.abs(p[k]).max()
            assert err < RTOL, (k, err)
        if s.opts.get("polar_palmo") == "on":
            err = np.abs(d["ef_induced_change"] - p["ef_induced_change"]).max() / np.abs(p["ef_induced"]).max()
            assert err < RTOL, ("ef_induced_change", err)
        if s.opts.get("polar_gs_ranked") == "on":
            assert np.array_equal(d["rank_metric"], p["rank_metric"])
    return o, p


@pytest.mark.parametrize("solver", ["SOLVER_GS_RANKED_PALMO", "SOLVER_JACOBI10"])
def test_config4_full_size_against_oracle(solver):
    """BASELINE config 4, N = 10 000 (8000 framework sites + 400 five-site H2), both solver variants of SURVEY 8d
    (src/System.Energy.cpp:3564-3598 contract_dipoles, :2534-2635 polar): cold, after a displace move, and restored."""
    from mpmcxx_b200 import workloads as W
    s = W.h2_framework(solver=getattr(W, solver))
    assert s.n == 10000 and int(np.count_nonzero(s.alpha)) == 9200      # 144 blocks of 64 (the last one 48 sites)
    e = _eng().Engine(s)
    o, _ = _compare_with_oracle(e, s)
    t = cases.displaced(s, seed=7)
    idx = np.nonzero(np.any(t.pos != s.pos, axis=1))[0]
    e.update_sites(int(idx[0]), t.pos[idx[0]:idx[-1] + 1])
    _compare_with_oracle(e, t)
    e.update_sites(int(idx[0]), s.pos[idx[0]:idx[-1] + 1])
    assert e.energy() == o                                               # restore(): bit-exact
    e.close()


def test_config4_two_kernel_pipeline_is_reproducible_and_agrees_with_single_launch_fallback():
    """The solver-cluster + updater-kernel pipeline gives the same BITS on every run at N = 10 000 — every row receives its panels
    in a fixed order whatever the timing — and agrees with the single-launch fallback (MPMC_GS_FUSED=1, what ncu profiles) to
    rounding: the fallback runs on a different grid, and the 60 chunks of rows the pipeline splits by columns over four warps
    (kernels_gs.cuh, gs_updater_body) add their column sums in a different association there.  With that split switched off the two
    are equal bit for bit."""
    from mpmcxx_b200 import workloads as W
    s = W.h2_framework(solver=W.SOLVER_GS_RANKED_PALMO)
    e = _eng().Engine(s)
    a, da = e.energy(), e.dipoles()
    e.close()
    e = _eng().Engine(s)
    a2, da2 = e.energy(), e.dipoles()
    a3 = e.energy()
    e.close()
    assert a == a2 == a3
    for k in ("mu", "ef_induced", "ef_induced_change"):
        assert np.array_equal(da[k], da2[k]), k
    os.environ["MPMC_GS_FUSED"] = "1"
    try:
        f = _eng().Engine(s)
        b, db = f.energy(), f.dipoles()
        f.close()
    finally:
        del os.environ["MPMC_GS_FUSED"]
    # with the column split switched off (developer switch 8) every row is summed exactly as in the fallback: the two-kernel
    # pipeline — solver cluster, updater kernel, flags across kernels — must then give the fallback's bits
    e = _eng().Engine(s)
    _eng()._ck(_eng().lib().mpmc_debug_gs_profile(e.h, 8, None, 0, None))
    c, dc = e.energy(), e.dipoles()
    e.close()
    assert c == b
    for k in ("mu", "ef_induced", "ef_induced_change"):
        assert np.array_equal(dc[k], db[k]), k
    assert a["polarization_iterations"] == b["polarization_iterations"] and a["iterator_failed"] == b["iterator_failed"]
    for k in ("polarization_energy", "energy"):
        assert abs(a[k] - b[k]) <= 1e-12 * abs(b[k]), k
    for k in ("mu", "ef_induced", "ef_induced_change"):
        scale = np.abs(db[k]).max()
        assert np.abs(da[k] - db[k]).max() <= 1e-12 * scale, k


def test_gauss_seidel_several_split_chunks_per_updater_cta_against_oracle():
    """10 100 polarizable sites = 2525 four-row chunks for the updaters' 2240 warps: 285 surplus chunks, i.e. up to three per CTA split
    by columns over groups of four warps (config 4 has at most one): every sub-term, the iteration count and all dipoles against
    the oracle."""
    from mpmcxx_b200 import workloads as W
    s = W.h2_framework(n_h2=700, solver=W.SOLVER_GS_RANKED_PALMO)
    assert int((s.alpha != 0).sum()) == 10100
    e = _eng().Engine(s)
    _compare_with_oracle(e, s)
    e.close()


@pytest.mark.parametrize("ncell,n_h2", [(12, 86), (10, 123)])
def test_gauss_seidel_tens_of_blocks_against_oracle(ncell, n_h2):
    """1728 + 258 = 1986 polarizable sites (31.03 blocks: a last block of two sites) and 1000 + 369 = 1369 (21.4 blocks)."""
    from mpmcxx_b200 import workloads as W
    s = W.h2_framework(ncell=ncell, n_h2=n_h2, solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt", seed=ncell * 1000 + n_h2)
    e = _eng().Engine(s)
    _compare_with_oracle(e, s)
    e.close()


def test_gauss_seidel_plain_and_precision_many_blocks():
    """polar_gs (list order in every sweep) with a fixed count, and with polar_precision (one launch per sweep, host decision)."""
    from mpmcxx_b200 import workloads as W
    for solver in ({"polar_gs": "on", "polar_max_iter": "5"}, W.SOLVER_GS_PRECISION):
        s = W.h2_framework(ncell=10, n_h2=77, solver=solver, ensemble="nvt", seed=31)
        e = _eng().Engine(s)
        _compare_with_oracle(e, s)
        e.close()


def test_config3_full_size_against_oracle_and_known_answer():
    """BASELINE config 3, N = 4096 LJ argon: the jittered start configuration against the oracle, and SURVEY 8c's known answer on
    the un-jittered 16^3 lattice (pairs exactly on the cutoff), -2709273.4197072242 from the unmodified reference."""
    from mpmcxx_b200 import workloads as W
    s = W.lj_argon()
    e = _eng().Engine(s)
    _compare_with_oracle(e, s, sites=False)
    e.close()
    k = W.lj_lattice(16, 60.0)
    e = _eng().Engine(k)
    o = e.energy()
    assert _rel(o["rd_energy"], -2709273.4197072242) < RTOL
    assert o["n_pairs_in_cutoff"] == 4311040.0
    e.close()


@pytest.mark.parametrize("five_site", [False, True])
def test_config5_all_beads_against_oracle(five_site):
    """BASELINE config 5 at full size: 512 molecules x 64 beads (single-site rd_only, and five-site + Ewald, N = 2560 per bead):
    the four per-bead energies of every bead, their means, the bead-spring chain and the kinetic term
    (src/SimulationControl.PathIntegral.cpp:752-828, 859-904)."""
    from mpmcxx_b200 import workloads as W
    from oracle import port
    tmpl, beads = W.pi_h2_cluster(n_side=8, P=64, five_site=five_site)
    e = _eng().Engine(tmpl, beads=beads)
    per, sums = e.pi_potential()
    outs = e.energy_all()
    q = port.pi_energy(tmpl, beads)
    P = beads.shape[0]
    scale = np.array([abs(o["es_real"]) + abs(o["es_self_intra"]) + abs(o["es_reciprocal"]) + abs(o["es_self"]) for o in outs])
    assert np.abs(per[:, 0] - q["per_bead"][:, 0]).max() / np.abs(q["per_bead"][:, 0]).max() < RTOL
    assert (np.abs(per[:, 1] - q["per_bead"][:, 1]) <= RTOL * np.maximum(scale, 1e-300)).all()
    assert _rel(sums[0] / P, q["rd"]) < RTOL
    assert abs(sums[1] / P - q["coulombic"]) <= RTOL * max(scale.mean(), 1e-300)
    chain, _, _ = e.pi_chain(closed=True)
    assert _rel(chain, q["chain_mass_len2"]) < 1e-12
    pot, means = e.pi_potential_allreduce(P)
    assert abs(pot - q["potential"]) <= RTOL * max(abs(q["rd"]) + scale.mean(), 1e-300)
    e.close()
