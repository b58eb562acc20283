"""The C-ABI library loads without a GPU and exports every symbol include/veros_b200.h declares.
No compute entry point is exercised beyond its descriptor validation (which runs before any CUDA call)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "veros_b200.h")


@pytest.fixture(scope="module")
def lib():
    from veros_b200 import _lib, build

    build.build()  # nvcc cross-compiles for sm_100a without a GPU; no-op when up to date
    return _lib.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(veros_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("veros_b200_solve_implicit_f64", "veros_b200_tdma_zmajor_f64", "veros_b200_tdma_zmajor_f32",
                 "veros_b200_iso_pre_f64", "veros_b200_iso_diffusion_f64", "veros_b200_iso_step_f64",
                 "veros_b200_last_error"):
        assert must in syms


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_python_binding_lists_match_header():
    from veros_b200 import _lib

    assert sorted(_lib.OPS + _lib.HELPERS) == declared_symbols()


def test_descriptor_layouts(lib):
    from veros_b200 import _lib

    assert lib.veros_b200_descriptor_size(0) == ctypes.sizeof(_lib.TridiagDescriptor) == 8  # reference: 2 ints
    assert lib.veros_b200_descriptor_size(1) == ctypes.sizeof(_lib.SolveDescriptor) == 16
    assert lib.veros_b200_descriptor_size(2) == ctypes.sizeof(_lib.IsoDescriptor) == 72
    assert lib.veros_b200_descriptor_size(3) == ctypes.sizeof(_lib.VmixDescriptor) == 24
    assert lib.veros_b200_abi_version() == _lib.ABI_VERSION


def test_bad_descriptor_is_latched_not_fatal(lib):
    from veros_b200 import _lib

    lib.veros_b200_clear_error()
    lib.veros_b200_iso_step_f64(None, None, b"short", 5)
    assert lib.veros_b200_last_error() == 100001
    assert b"bad descriptor" in lib.veros_b200_last_error_string()
    with pytest.raises(RuntimeError, match="bad descriptor"):
        _lib.check_error("test")
    assert lib.veros_b200_last_error() == 0  # check_error clears the latch
    bad = _lib.IsoDescriptor(nx_tot=9, ny_tot=9, nz=5, eq_of_state_type=7, iso_dslope=1.0, dt_tracer=1.0)
    lib.veros_b200_iso_pre_f64(None, None, bytes(bad), ctypes.sizeof(bad))
    assert lib.veros_b200_last_error() == 100002
    lib.veros_b200_clear_error()


def test_workspace_sizes(lib):
    from veros_b200 import _lib

    d = _lib.IsoDescriptor(nx_tot=20, ny_tot=10, nz=7, eq_of_state_type=1, iso_dslope=1e-3, iso_slopec=1e-3, dt_tracer=1.0)
    n3 = 20 * 10 * 7  # even, so no alignment padding in the sizes below
    o = bytes(d)
    pre1 = lib.veros_b200_iso_pre_workspace_bytes(o, len(o))
    d.eq_of_state_type = 5
    o5 = bytes(d)
    pre5 = lib.veros_b200_iso_pre_workspace_bytes(o5, len(o5))
    assert pre1 > 0 and pre5 - pre1 == 2 * 8 * n3  # drdT, drdS scratch for TEOS-10
    # fluxes (3) + dissipation (1) per tracer, metric tables
    assert lib.veros_b200_iso_diffusion_workspace_bytes(o, len(o)) == 4 * 8 * n3 + pre1
    # the step's scratch covers whichever implementation the flags choose: separate launches need the full-size
    # flux / dissipation arrays, the fused persistent kernel a ring of planes plus its counters
    assert lib.veros_b200_iso_step_workspace_bytes(o5, len(o5)) >= 8 * 8 * n3 + pre5
    assert 0 < lib.veros_b200_iso_step_stats_offset(o5, len(o5)) < lib.veros_b200_iso_step_workspace_bytes(o5, len(o5))
    assert lib.veros_b200_iso_step_workspace_bytes(b"x", 1) == 0
    lib.veros_b200_clear_error()


def test_missing_library_fails_loudly(monkeypatch):
    from veros_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libveros_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "veros_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "oracle/"):
                    assert needle not in src, (fn, needle)
