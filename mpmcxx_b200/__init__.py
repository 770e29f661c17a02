"""mpmcxx_b200 — B200-native (sm_100a) energy engine for the mpmc++ Monte Carlo code.

The product is `libmpmc_b200.so` (hand-written CUDA kernels behind the C-ABI of include/mpmc_b200.h).  This package
holds its sources (csrc/), the in-tree build, a thin ctypes binding used by the tests and bench.py, and generators for
the synthetic systems the benchmark configs name.
"""
__all__ = ["build", "config", "engine", "workloads"]
