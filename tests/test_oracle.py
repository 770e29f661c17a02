"""Pin the CPU oracle (oracle/oracle.c) against the reference: committed golden vectors generated from the
unmodified reference (tests/golden/make_golden.py), and — where oracle/_ref is present — the reference itself."""
import os

import numpy as np
import pytest

from oracle import port, ref
from tests import cases

# per-term tolerances: everything is FP64 on both sides; only summation order differs
RTOL = 2e-11


def _rel(a, b, scale=None):
    s = max(abs(b), 1e-300) if scale is None else scale
    return abs(a - b) / s


def _check_terms(o, r):
    for k_o, k_r in (("rd", "ref_rd"), ("lrc_pair", "ref_lrc_pair"), ("lrc_self", "ref_lrc_self"), ("es_real", "ref_es_real"),
                     ("es_self_intra", "ref_es_self_intra"), ("es_recip", "ref_es_recip"), ("es_self", "ref_es_self"),
                     ("polar", "ref_polar")):
        a, b = o[k_o], float(r[k_r])
        if b == 0.0:
            assert abs(a) < 1e-9, (k_o, a, b)
        else:
            assert _rel(a, b) < RTOL, (k_o, a, b)
    # the Ewald total cancels ~1e5 K down to a few K: judge it against the sum of |sub-terms|
    tot_o = o["es_real"] - o["es_self_intra"] + o["es_recip"] + o["es_self"]
    tot_r = float(r["ref_es_real_minus_intra"]) + float(r["ref_es_recip"]) + float(r["ref_es_self"])
    scale = abs(o["es_real"]) + abs(o["es_self_intra"]) + abs(o["es_recip"]) + abs(o["es_self"])
    assert abs(tot_o - tot_r) <= RTOL * max(scale, 1e-300)
    assert int(o["iterations"]) == int(r["ref_iterations"])
    assert int(o["iterator_failed"]) == int(r["ref_iterator_failed"])


@pytest.mark.parametrize("name", sorted(cases.CLASSIC))
def test_port_matches_golden(name):
    s, r = cases.load_golden(name)
    o = port.energy(s)
    _check_terms(o, r)
    assert _rel(o["cutoff"], float(r["cell_cutoff"])) < 1e-15
    assert _rel(o["ewald_alpha"], float(r["cell_ewald_alpha"])) < 1e-15
    if int(r["ref_iterations"]) or s.opts.get("polar_zodid") == "on":
        for k in ("mu", "ef_static", "ef_induced", "ef_induced_change"):
            ref_v = r["ref_" + k]
            scale = max(np.abs(ref_v).max(), 1e-300)
            # ef_induced_change under Jacobi is pure rounding noise in the reference (it cancels analytically)
            tol = 1e-9 if k != "ef_induced_change" else 1e-6
            assert np.abs(o[k] - ref_v).max() / scale < tol, k
        if s.opts.get("polar_gs_ranked") == "on":
            assert np.array_equal(o["rank_metric"], r["ref_rank_metric"])
    # energy() as the MC loop sees it: cold total and the total after a one-molecule move
    e_cold = r["ref_energy_cold"]
    assert _rel(o["energy"], e_cold[0], scale=abs(e_cold[1]) + abs(e_cold[2]) + abs(e_cold[3])) < 1e-10
    o2 = port.energy(s, pos=r["moved_pos"], want_sites=False)
    e_mv = r["ref_energy_moved"]
    assert _rel(o2["rd_total"], e_mv[1]) < RTOL
    assert _rel(o2["polar"], e_mv[3]) < RTOL or e_mv[3] == 0.0


@pytest.mark.parametrize("name", sorted(cases.PI))
def test_port_pi_matches_golden(name):
    s, r = cases.load_golden(name)
    o = port.pi_energy(s, r["beads"])
    assert _rel(o["rd"], float(r["ref_pi_rd"])) < RTOL
    assert _rel(o["chain_mass_len2"], float(r["ref_pi_chain_mass_len2"])) < 1e-13
    assert _rel(o["kinetic"], float(r["ref_pi_kinetic"])) < 1e-12
    assert abs(o["potential"] - float(r["ref_pi_potential"])) < 1e-9 * max(1.0, abs(float(r["ref_pi_potential"])))
    if float(r["ref_pi_polar"]) != 0.0:
        assert _rel(o["polar"], float(r["ref_pi_polar"])) < RTOL


def test_survey_known_answers():
    """SURVEY.md §8c numbers, independent of our fixtures."""
    s, _ = cases.load_golden("lj_lattice_8")
    assert abs(port.energy(s)["rd_total"] - (-338477.27924728388)) < 1e-6
    s, _ = cases.load_golden("kat_jacobi10")
    o = port.energy(s)
    assert abs(o["rd_total"] - (-2349.4779825087139)) < 1e-8
    assert abs(o["coulombic"] - 4.0111619445960969) < 1e-6
    assert abs(o["polar"] - (-8.4466089936124327)) < 1e-10


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["kat_gs_ranked_palmo", "tri_gs_ranked_palmo", "h2fw_6_jacobi10"])
def test_port_matches_live_reference(name):
    s = cases.CLASSIC[name]()
    r = ref.RefSystem(s, ensemble="nvt")
    t = r.terms()
    o = port.energy(s)
    for k in ("rd", "lrc_pair", "es_real", "es_self_intra", "es_recip", "es_self", "polar"):
        assert _rel(o[k], t[k]) < RTOL, k
    c = r.cell()
    pc = port.cell(s.basis)
    assert np.allclose(pc["recip"], c["recip"], rtol=0, atol=0)
    assert pc["cutoff"] == c["cutoff"] and pc["volume"] == c["volume"]


def test_root_averages_restate_the_reference():
    """mpmcxx_b200/averages.py root_averages == System::update_root_averages of the unmodified reference (tests/golden/root_averages.npz,
    made by `make_golden.py rootavg`) on 300 samples of a uVT job with a frozen framework: means, errors, density, heat capacity,
    compressibility, weight percent, excess adsorption, pore density, qst — the same arithmetic, to the last bits."""
    from mpmcxx_b200 import averages
    from tests import cases
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "root_averages.npz"))
    s = cases.uvt_pore_for_averages()
    x = cases.root_average_samples()
    assert np.array_equal(x, z["samples"])
    mobile = ~s.frozen.astype(bool)
    particle_mass = float(s.mass[s.mol == s.mol[mobile][0]].sum())
    a = averages.root_averages(x, float(s.opts["temperature"]), float(z["volume"]), particle_mass, frozen_mass=float(z["frozen_mass"]),
                               free_volume=float(s.opts["free_volume"]), fugacity=float(z["fugacity"]))
    assert float(z["fugacity"]) > 0
    for k, v in zip(z["keys"], z["values"]):
        assert abs(a[str(k)] - v) <= 1e-13 * abs(v), (str(k), a[str(k)], float(v))
    # the two-column helper used by the comparison tests is the same recursion
    m, e = averages.root_average(x[:, 0])
    assert m == a["energy"] and abs(e - a["energy_error"]) <= 1e-15 * e


def test_merged_chains_are_the_root_averages_of_the_interleaved_samples():
    """Independent chains are merged as the reference's head node merges MPI ranks (src/System.MonteCarlo.cpp:1973-2022): node after
    node at every sample time.  With one chain it is root_averages; with K equal chains the means stay and the errors shrink."""
    from mpmcxx_b200 import averages
    from tests import cases
    x = cases.root_average_samples(n=120, seed=5)
    sysd = dict(temperature=77.0, volume=8000.0, particle_mass=2.016, frozen_mass=960.0, free_volume=8000.0)
    one = averages.root_averages(x, **sysd)
    assert averages.merge_chains([x], **sysd) == one
    four = averages.merge_chains([x, x, x, x], **sysd)
    assert abs(four["energy"] - one["energy"]) < 1e-12 * abs(one["energy"]) and abs(four["N"] - one["N"]) < 1e-12 * one["N"]
    assert 0.45 < four["energy_error"] / one["energy_error"] < 0.55          # sqrt((n - 1) / (4 n - 1)) ~ 1/2
    y = cases.root_average_samples(n=120, seed=6)
    two = averages.merge_chains([x, y], **sysd)
    inter = np.empty((240, 6)); inter[0::2] = x; inter[1::2] = y
    assert two == averages.root_averages(inter, **sysd)
