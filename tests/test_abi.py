"""CPU-side checks of the drop-in boundary: the shared library builds, loads, and exports every symbol that
include/mpmc_b200.h declares; without a GPU the entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

from mpmcxx_b200 import build, config, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(engine.EXPORTS) == names
    assert L.mpmc_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes follow from the header's field lists: 9 doubles + 6 ints(+pad) ... keep the ctypes mirror honest
    assert C.sizeof(engine.MpmcConfig) % 8 == 0
    assert C.sizeof(engine.MpmcEnergyOut) == 15 * 8 + 6 * 4
    assert engine.MpmcConfig.ewald_alpha.offset % 8 == 0 and engine.MpmcConfig.polar_damp.offset % 8 == 0


def test_no_cpu_fallback_and_validation():
    import numpy as np
    L = engine.lib()
    n = C.c_int(-1)
    rc = L.mpmc_device_count(C.byref(n))
    o = config.EnergyOptions()
    cfg = engine.make_config(np.eye(3) * 20.0, o)
    h = C.c_void_p()
    rc2 = L.mpmc_create(C.byref(cfg), C.byref(h))
    if rc != 0 or n.value == 0:
        assert rc2 == 30000 and not h.value          # MPMC_ERR_CUDA: loud failure, nothing created
        assert L.mpmc_last_error()
    else:
        assert rc2 == 0
        L.mpmc_destroy(h)
    # validation happens before any device work (reference: SimulationControl.cpp:2612-2700)
    o2 = config.EnergyOptions(polarization=1, polar_iterative=0, damp_type=2, polar_damp=2.1)
    rc3 = L.mpmc_create(C.byref(engine.make_config(np.eye(3) * 20.0, o2)), C.byref(h))
    assert rc3 == 4004
    o3 = config.EnergyOptions(polarization=1, polar_iterative=1, damp_type=2, polar_damp=2.1, polar_precision=1e-6, polar_max_iter=4)
    assert L.mpmc_create(C.byref(engine.make_config(np.eye(3) * 20.0, o3)), C.byref(h)) == 4002


def test_sources_do_not_use_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    bad = re.compile(r"(^|\n)\s*(from|import)\s+oracle\b|liboracle|libmpmc_ref|#include\s+\"[^\"]*oracle")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpmcxx_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)
    assert os.path.exists(build.LIB)


def test_product_libraries_do_not_contain_or_link_the_oracle():
    """The oracle is test infrastructure: neither shipped library may define, import or depend on anything of it."""
    import shutil
    import subprocess
    if not shutil.which("nm") or not shutil.which("readelf"):
        import pytest
        pytest.skip("binutils not available")
    for lib in (build.build_library(), build.build_host()):
        syms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True, check=True).stdout
        assert not re.search(r"\borc_[a-z_]+", syms), lib
        needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True, check=True).stdout
        assert "oracle" not in needed and "mpmc_ref" not in needed, lib
    for root, _, files in os.walk(os.path.join(ROOT, "mpmcxx_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle.h" not in text and "liboracle" not in text and not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
