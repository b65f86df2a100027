"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/osph.h declares.
No compute call is made here (there is no GPU in the build container and the library has no CPU path)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from osph_b200 import capi


@pytest.fixture(scope="module")
def libpath():
    return capi.build()


def test_header_symbols_exported(libpath):
    hdr = open(os.path.join(ROOT, "include", "osph.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(osph_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = C.CDLL(libpath)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    bound = set(capi.lib()._osph_signatures)
    assert declared == bound, declared ^ bound


def test_config_layout_and_defaults(libpath):
    cfg = capi.default_config(25.0, 0.5, 1000.0)
    assert cfg.struct_size == C.sizeof(capi.Config)
    assert cfg.co == 221.47234590350104 and cfg.B == 7007142.857142859 and cfg.D == 1226.25
    assert (cfg.gamma, cfg.alpha, cfg.beta, cfg.epsilon, cfg.p1, cfg.p2) == (7.0, 0.01, 0.0, 0.5, 4.0, 2.0)
    assert (cfg.nn_scale, cfg.h_sigma, cfg.cfl_courant, cfg.cfl_force, cfg.gravity) == (2.0, 1.3, 0.25, 0.25, 9.81)
    assert capi.lib().osph_version() == 1


def test_no_cpu_fallback(libpath):
    """Without a CUDA device context creation must fail loudly; with one it must succeed."""
    import torch
    cfg = capi.default_config(25.0, 0.5)
    if torch.cuda.is_available():
        capi.Context(cfg).close()
    else:
        with pytest.raises(capi.OsphError) as e:
            capi.Context(cfg)
        assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_leaf_equations_have_no_cpu_path_either(libpath):
    """The stand-alone equations (src.Equations.*) validate their arguments on the host and otherwise need the device:
    without one they raise, they never compute on the CPU."""
    import numpy as np
    import torch
    L = capi.lib()
    out = np.full(7, 3.0)
    dp = C.POINTER(C.c_double)
    # J < 0, or J > 0 without columns: invalid, before any CUDA call
    assert L.osph_leaf_equations(0, -1, None, None, 0, 1, 1, 0, 0, 0, 0, 0, 0, 4, 2, out.ctypes.data_as(dp)) == -1
    assert L.osph_leaf_equations(0, 5, None, None, 0, 1, 1, 0, 0, 0, 0, 0, 0, 4, 2, out.ctypes.data_as(dp)) == -1
    # an empty table is all zeros and needs no device
    assert L.osph_leaf_equations(0, 0, None, None, 0, 1, 1, 0, 0, 0, 0, 0, 0, 4, 2, out.ctypes.data_as(dp)) == 0
    assert not out.any()
    if not torch.cuda.is_available():
        from src.Common import computed_dtype
        from src.Equations.Continuity import Continuity
        from src.Equations.Courant import Courant
        comp = np.zeros(4, dtype=computed_dtype)
        with pytest.raises(capi.OsphError):
            Continuity(np.array([]), comp)
        with pytest.raises(capi.OsphError):
            Courant(0.4, np.ones(3), np.ones(3))


def test_bad_struct_size_rejected(libpath):
    cfg = capi.default_config(25.0, 0.5)
    cfg.struct_size = 8
    with pytest.raises(capi.OsphError) as e:
        capi.Context(cfg)
    assert e.value.code == -1


def test_product_never_touches_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "offshore-sph_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', 'Makefile')):
                txt = open(os.path.join(d, f), errors='ignore').read()
                if re.search(r"\boracle\b", txt) and 'oracle' in txt.replace("CPU oracle", ""):
                    for line in txt.splitlines():
                        if re.search(r"^\s*(from|import)\s+oracle|liboracle|oracle/", line):
                            bad.append((f, line.strip()))
    assert not bad, bad


def test_product_never_touches_the_emulator():
    """tests/emu (the SIMT-emulated CPU build of the kernel sources) is test infrastructure: nothing in the product
    package, in bench.py or in __graft_entry__.py may name it, and the default library path is the nvcc build."""
    bad = []
    roots = [os.path.join(ROOT, "offshore-sph_b200"), os.path.join(ROOT, "include")]
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for r in roots:
        for d, _, fs in os.walk(r):
            files += [os.path.join(d, f) for f in fs if f.endswith(('.py', '.cu', '.cuh', '.h', 'Makefile'))]
    for f in files:
        txt = open(f, errors='ignore').read()
        if re.search(r"libosph_emu|tests/emu|OSPH_EMU\b|emu::", txt):
            bad.append(f)
    assert not bad, bad
    from osph_b200 import capi
    if not os.environ.get("OSPH_LIB"):
        assert capi.LIB_PATH.endswith(os.path.join("lib", "libosph_b200.so"))
