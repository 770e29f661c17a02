"""Ensemble averages of a Markov chain from its step log, the way the reference accumulates them
(src/System.Averages.cpp:8-208 update_root_averages: running mean, mean of squares, error = sqrt(<x^2> - <x>^2) / sqrt(m - 1) over the
samples taken every `corrtime` steps; src/System.MonteCarlo.cpp:1973-2022 merges independent chains by the same formula), plus a
blocked standard error that does not assume uncorrelated samples — what a comparison "within statistical error" needs.

Host-side bookkeeping only (numpy).  The step log is what mpmc_host_run / the reference harness return per step:
[move type, trial energy (classic) or trial potential (pi_nvt), Boltzmann factor, accepted, N (classic) or kinetic energy (pi_nvt)].
"""
from __future__ import annotations

import numpy as np


def chain_series(log: np.ndarray, first: float) -> dict:
    """Observables of the CURRENT state after every step: a rejected move keeps the previous state's energy.  `first` is the energy
    (potential for pi_nvt) of the start configuration.  Column 4 of the log already belongs to the current state."""
    e = np.empty(len(log))
    cur = first
    for i in range(len(log)):
        if log[i, 3]:
            cur = log[i, 1]
        e[i] = cur
    return {"energy": e, "aux": log[:, 4].copy(), "accepted": log[:, 3].copy()}


def root_average(samples: np.ndarray) -> tuple:
    """(mean, error) exactly as update_root_averages accumulates them (Averages.cpp:28-40)."""
    avg = avg_sq = 0.0
    for k, x in enumerate(samples, start=1):
        f = (k - 1.0) / k
        avg = f * avg + x / k
        avg_sq = f * avg_sq + x * x / k
    m = len(samples)
    err = np.sqrt(max(avg_sq - avg * avg, 0.0)) / np.sqrt(m - 1.0) if m > 1 else float("nan")
    return avg, err


# src/constants.h
_KB, _NA, _A32CM3, _ATM2PASCALS, _METER2ANGSTROM, _ATM2REDUCED = 1.3806503e-23, 6.0221415e23, 1.0e-24, 101325.0, 1.0e10, 0.0073389366


def root_averages(samples: np.ndarray, temperature: float, volume: float, particle_mass: float, frozen_mass: float = 0.0,
                  free_volume: float = 0.0, fugacity: float = 0.0) -> dict:
    """update_root_averages (src/System.Averages.cpp:8-208) over `samples` [n, 6] = (energy, coulombic, rd, polarization, N, NU) of a
    constant-volume ensemble: running means and errors of the energies and of N, and what the reference derives from them —
    density (g/cm^3), heat capacity and compressibility with their errors (the Stirling form of the gamma ratio, :147-160), weight
    percent, excess adsorption (mg/g), pore density and the isosteric heat qst (kJ/mol) when the system has a frozen framework.
    `fugacity` is the reference's fugacities[0]: the value a fugacity keyword (h2_fugacity, user_fugacities ...) put there, 0 otherwise
    — `fugacities` is an array member, so the branch that would use the PRESSURE instead (:189-194) is never taken."""
    x = np.asarray(samples, dtype=np.float64)
    names = ("energy", "coulombic_energy", "rd_energy", "polarization_energy", "N")
    avg = {k: 0.0 for k in names}
    sq = {k: 0.0 for k in names}
    err = {k: float("nan") for k in names}
    nu = dens = dens_sq = 0.0
    out = {}
    for i in range(len(x)):
        m = float(i + 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            sdom = np.float64(1.0) / np.sqrt(np.float64(m - 1.0))
        f = (m - 1.0) / m
        for c, k in enumerate(names):
            v = x[i, c] if c < 4 else x[i, 4]
            avg[k] = f * avg[k] + v / m
            sq[k] = f * sq[k] + (v * v) / m
            with np.errstate(invalid="ignore"):
                err[k] = sdom * np.sqrt(np.float64(sq[k] - avg[k] * avg[k]))
        nu = f * nu + x[i, 5] / m
        cur = x[i, 4] * particle_mass / (volume * _NA * _A32CM3)
        dens = f * dens + cur / m
        dens_sq = f * dens_sq + cur * cur / m
        with np.errstate(invalid="ignore", divide="ignore"):
            dens_err = sdom * np.sqrt(np.float64(dens_sq - dens * dens))
            g = np.power(np.float64(m - 2.0) / np.float64(m - 1.0), 0.5 * m - 1.0) * np.sqrt(np.float64(0.5 * (m - 2.0))) * np.exp(0.5)
            g = np.sqrt(np.float64(1.0 / m * (m - 1.0 - 2.0 * g * g)))
            hc = (_KB * _NA / 1000.0) * (sq["energy"] - avg["energy"] * avg["energy"]) / (temperature * temperature)
            comp = _ATM2PASCALS * (volume / _METER2ANGSTROM ** 3) * (sq["N"] - avg["N"] * avg["N"]) / (_KB * temperature * avg["N"] * avg["N"])
            out = {"density": dens, "density_error": dens_err, "heat_capacity": hc, "heat_capacity_error": sdom * 2.0 * g * hc,
                   "compressibility": comp, "compressibility_error": sdom * 2.0 * g * comp, "NU": nu}
            if frozen_mass > 0.0:
                out["percent_wt"] = 100.0 * avg["N"] * particle_mass / (frozen_mass + avg["N"] * particle_mass)
                out["percent_wt_me"] = 100.0 * avg["N"] * particle_mass / frozen_mass
                if free_volume > 0.0:
                    out["excess_ratio"] = 1000.0 * (avg["N"] * particle_mass - (particle_mass * free_volume * fugacity * _ATM2REDUCED) / temperature) / frozen_mass
                    out["pore_density"] = cur * volume / free_volume
                q = -(nu - avg["N"] * avg["energy"])
                q /= (sq["N"] - avg["N"] * avg["N"])
                q += temperature
                out["qst"] = q * _KB * _NA / 1000.0
    for k in names:
        out[k] = avg[k]
        out[k + "_error"] = float(err[k])
    return out


def merge_chains(chains, **system) -> dict:
    """Independent chains merged the way the reference's head node does it with MPI ranks (src/System.MonteCarlo.cpp:1973-2022): at
    every correlation time the observables of node 0, 1, ... are averaged in one after the other — so the root averages run over
    samples in the order (time, node) and the error bars shrink with sqrt(nodes x samples).  `chains`: one [n, 6] sample array per
    chain (same n), as root_averages takes; the remaining arguments are root_averages's."""
    x = np.stack([np.asarray(c, dtype=np.float64) for c in chains], axis=1)      # [time, node, 6]
    return root_averages(x.reshape(-1, x.shape[-1]), **system)


def block_means(series: np.ndarray, nblocks: int, discard: float = 0.25) -> np.ndarray:
    """Means over `nblocks` consecutive blocks after dropping the first `discard` fraction (equilibration)."""
    s = series[int(len(series) * discard):]
    n = len(s) // nblocks
    return s[: n * nblocks].reshape(nblocks, n).mean(axis=1)


def compare(blocks_a: np.ndarray, blocks_b: np.ndarray) -> dict:
    """Two sets of block means (any number of chains each, flattened): difference of the means in units of the combined standard
    error of the means."""
    a, b = np.ravel(blocks_a), np.ravel(blocks_b)
    ma, mb = a.mean(), b.mean()
    sa, sb = a.std(ddof=1) / np.sqrt(len(a)), b.std(ddof=1) / np.sqrt(len(b))
    sig = float(np.hypot(sa, sb))
    return {"mean_a": float(ma), "mean_b": float(mb), "sem_a": float(sa), "sem_b": float(sb), "sigma": sig,
            "z": float(abs(ma - mb) / sig) if sig > 0 else (0.0 if ma == mb else float("inf"))}
