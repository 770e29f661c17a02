"""Build the engine's shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpmc_b200.so")
HOSTLIB = os.path.join(HERE, "libmpmc_host.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]
NVCC_LIBS = ["-ldl"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without the CUDA toolkit")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def engine_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "mpmc_b200.h"))
    return srcs


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> mpmcxx_b200/libmpmc_b200.so"""
    if force or _stale(LIB, engine_sources()):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "engine.cu")] + NVCC_LIBS
        subprocess.run(cmd, check=True)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_BIN = os.path.join(HERE, "mpmcxx-b200")
HOST_FLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-Wall"]


def host_sources():
    return [os.path.join(HOST_DIR, f) for f in sorted(os.listdir(HOST_DIR)) if f.endswith((".cpp", ".h"))] + [os.path.join(ROOT, "include", "mpmc_b200.h")]


def build_host(force: bool = False) -> str:
    """g++ -> mpmcxx_b200/libmpmc_host.so (the C++ mirror of System/Molecule/Atom/SimulationControl over the C-ABI) and the
    mpmcxx-b200 command-line driver; both link libmpmc_b200.so through an $ORIGIN rpath."""
    build_library(force=False)
    if force or _stale(HOSTLIB, host_sources()) or _stale(HOST_BIN, host_sources()):
        cxx = "g++"
        srcs = [os.path.join(HOST_DIR, f) for f in ("host.cpp", "sim_control.cpp", "host_capi.cpp")]
        link = ["-L" + HERE, "-lmpmc_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.run([cxx] + HOST_FLAGS + ["-shared", "-o", HOSTLIB] + srcs + link, check=True)
        subprocess.run([cxx] + HOST_FLAGS + ["-o", HOST_BIN, os.path.join(HOST_DIR, "main.cpp")] + srcs + link, check=True)
    return HOSTLIB
