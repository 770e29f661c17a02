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
