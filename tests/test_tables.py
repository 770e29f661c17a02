"""Host-side numerics of the CUDA path that need no GPU: the r^2-indexed radial tables (radial_table.h) against scipy /
mpmath, and the cutoff thresholds on r^2 against the reference's own tests (System.Energy.cpp:934, :1490)."""
import numpy as np
import pytest
from scipy import special

from mpmcxx_b200 import engine


@pytest.mark.parametrize("cutoff", [8.0, 30.0, 40.0, 50.0])
def test_erfc_table_matches_libm(cutoff):
    alpha = 3.5 / cutoff
    rs = np.random.RandomState(3)
    u = np.exp(rs.uniform(np.log(0.26), np.log(cutoff ** 2), 200000))
    t = engine.radial_table(0, alpha, 0.25, cutoff ** 2 * 1.001, u)
    r = np.sqrt(u)
    ref = special.erfc(alpha * r) / r
    # absolute error on the near-field scale (values of order 1): what a sum over pairs sees; libm's own erfc is good to ~1e-15
    assert np.max(np.abs(t - ref)) < 1e-15
    near = u < (0.5 * cutoff) ** 2
    assert np.max(np.abs(t[near] - ref[near]) / ref[near]) < 5e-15


def test_erfc_table_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    alpha = 3.5 / 40.0
    u = np.array([0.3, 1.0, 7.77, 16.0, 123.456, 900.0, 1599.0])
    t = engine.radial_table(0, alpha, 0.25, 1601.0, u)
    for ui, ti in zip(u, t):
        r = mp.sqrt(mp.mpf(float(ui)))
        ex = mp.erfc(mp.mpf(alpha) * r) / r
        # one ulp of the value, except in the last octave before the cutoff where the value itself is 1e-8 of the near field
        # and the table is held to an absolute error far below anything a sum can see
        if ui <= 900.0:
            assert abs(mp.mpf(float(ti)) - ex) / ex < mp.mpf(5e-16)
        else:
            assert abs(mp.mpf(float(ti)) - ex) < mp.mpf(1e-20)


def test_field_table_matches_closed_form():
    a = 3.5 / 40.0
    rs = np.random.RandomState(5)
    u = np.exp(rs.uniform(np.log(0.26), np.log(1600.0), 100000))
    f0, f1 = engine.radial_table(1, a, 0.25, 1601.6, u)
    r = np.sqrt(u)
    g = 2.0 * a / np.sqrt(np.pi) * np.exp(-a * a * u) * r
    ref0 = (g + special.erfc(a * r)) / (u * r)
    near = u < 400.0
    assert np.max(np.abs(f0[near] - ref0[near]) / ref0[near]) < 2e-13           # 16 intervals per octave: ~2e-14 on a power law
    assert np.max(np.abs(f0[~near] - ref0[~near])) < 1e-16                      # beyond, the factor is < 1e-6 of the near field
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    for ui, v in list(zip(u, f1))[:100]:         # the excluded form cancels in double arithmetic: compare with exact arithmetic
        rr, am, um = mp.sqrt(mp.mpf(float(ui))), mp.mpf(a), mp.mpf(float(ui))
        ex = (2 * am / mp.sqrt(mp.pi) * mp.e ** (-am ** 2 * um) * rr - mp.erf(am * rr)) / (um * rr)
        assert abs((mp.mpf(float(v)) - ex) / ex) < mp.mpf(1e-13)


@pytest.mark.parametrize("cutoff", [10.0, 30.0, 40.0, 12.345678901234567, 50.0])
def test_cutoff_thresholds_are_the_reference_tests(cutoff):
    t2_lj, t2_es = engine.cutoff_thresholds(cutoff)
    small = 1.0e-12                                    # constants.h:54
    lj = lambda x: np.sqrt(x) - small < cutoff         # System.Energy.cpp:934
    es = lambda x: not (np.sqrt(x) > cutoff)           # :1490
    for t2, pred in ((t2_lj, lj), (t2_es, es)):
        assert pred(t2) and not pred(np.nextafter(t2, np.inf))
        for k in range(1, 50):                         # monotone around the threshold
            assert pred(np.float64(t2) * (1 - k * 1e-16)) and not pred(np.float64(t2) * (1 + k * 4e-16))
    assert t2_lj >= t2_es
