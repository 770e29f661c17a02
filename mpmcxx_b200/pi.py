"""Host-side bookkeeping for path-integral beads sharded over ranks (one process per GPU).

The reference runs one MPI rank per bead, every rank holding all P geometries and replaying identical moves from a broadcast
seed, and exchanges four scalars per energy evaluation (src/SimulationControl.PathIntegral.cpp:757-766,
src/SimulationControl.cpp:53-65,98).  Here a rank owns a contiguous slice of beads; the exchange is one all-reduce (sum) of the
four per-bead-sum scalars, and PI_calculate_potential's means (:798-801) are taken over the GLOBAL Trotter number.
"""
from __future__ import annotations

import numpy as np

# constants.h:15-24
KB = 1.3806503e-23
H = 6.626068e-34
HBAR2 = 1.11211999e-68
AMU2KG = 1.66053873e-27


def bead_range(P: int, rank: int, world: int):
    """Contiguous slice of bead indices owned by `rank`; P must divide evenly (P is a power of two >= 4 in the reference,
    PathIntegral.cpp:552-580, and so is the GPU count of one box)."""
    if P % world:
        raise ValueError("Trotter number %d is not divisible by %d ranks" % (P, world))
    per = P // world
    return rank * per, (rank + 1) * per


def combine_potential(local_sums, P: int, all_reduce=None):
    """local_sums: this rank's sums over its beads of (rd, coulombic, polarization, vdw).  all_reduce(array) -> summed in place.
    Returns (potential, means[4]) exactly as PI_calculate_potential aggregates (:786-804)."""
    s = np.array(local_sums, dtype=np.float64)
    if all_reduce is not None:
        all_reduce(s)
    means = s / P
    return float(means[0] + means[1] + means[3] + means[2]), means


def chain_mass_len2(coms: np.ndarray, mass_amu: float) -> float:
    """PI_chain_mass_length2 (:916-970) of one molecule: coms is (P, 3), ring closed."""
    d = coms - np.roll(coms, -1, axis=0)
    return float((d * d).sum() * (mass_amu * AMU2KG) * 1e-20)


def kinetic_estimator(chain_total: float, n_mobile: int, P: int, T: float) -> float:
    """PI_calculate_kinetic (:810-828), in Kelvin."""
    beta = 1.0 / (KB * T)
    omega2 = P / (beta * beta * HBAR2)
    return (1.0 / KB) * (0.5 * 3.0 * n_mobile * KB * T * P - 0.5 * omega2 * chain_total)


def bead_perturb_boltzmann(d_potential: float, d_chain: float, P: int, T: float) -> float:
    """PI_NVT_boltzmann_factor, MOVETYPE_PERTURB_BEADS branch (:503-523) without the orientational term."""
    k = (P * np.pi * np.pi * KB * T) / (2.0 * H * H)
    return float(np.exp(min(700.0, -d_potential / T - d_chain * k)))
