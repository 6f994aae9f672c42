"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/galax_b200.h declares;
host-side marshalling and argument checking (no compute calls -- there is no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "galax_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built_lib):
    from galax_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in galax_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(syms)
    assert built_lib.gx_version() == 100
    assert built_lib.gx_strerror(-2).decode().startswith("unsupported")
    assert built_lib.gx_workspace_bytes() >= 8


def test_library_is_sm100a_native(built_lib):
    import subprocess

    from galax_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layout_matches_header():
    from galax_b200 import _lib

    assert C.sizeof(_lib.GxComponent) == 8 + 64 + 64  # kind, reserved, p[8], dp[8]
    assert C.sizeof(_lib.GxPotential) == 16 + 136 * _lib.GX_MAX_COMPONENTS
    assert C.sizeof(_lib.GxPid) == 8 * 10 + 8 + 8


def test_argument_errors_without_gpu(built_lib):
    """Argument validation happens before any CUDA call."""
    from galax_b200 import _lib

    P = _lib.GxPotential()
    P.n = 1
    P.G = 1.0
    P.c[0].kind = 99
    assert built_lib.gx_potential_eval(C.byref(P), None, 0.0, 0, 1, None, None, None, None, None) == -2
    P.c[0].kind = _lib.KIND_PLC
    P.c[0].p[0], P.c[0].p[1], P.c[0].p[2] = 1.0, 2.5, 1.0  # alpha >= 2 unsupported
    assert built_lib.gx_potential_eval(C.byref(P), None, 0.0, 0, 1, None, None, None, None, None) == -2
    P.c[0].kind = _lib.KIND_HERNQUIST
    assert built_lib.gx_potential_eval(C.byref(P), None, 0.0, 5, 1, None, None, None, None, None) == -1
    assert built_lib.gx_integrate_fixed(C.byref(P), None, None, 0, 0.0, 1.0, -0.1, None, 0, 0, -1, 0, None, None,
                                        None, None) == -1  # dt0 against the direction of integration
    assert built_lib.gx_integrate_fixed(C.byref(P), None, None, 0, 0.0, 1.0, 0.1, None, 0, 7, -1, 0, None, None,
                                        None, None) == -1  # unknown scheme
    P.n = 99
    assert built_lib.gx_potential_eval(C.byref(P), None, 0.0, 0, 1, None, None, None, None, None) == -1


def test_host_mirror_marshalling():
    import galax_b200.potential as gp
    from oracle import potentials as op

    for cls, ofun in ((gp.MilkyWayPotential, op.milky_way_potential), (gp.MilkyWayPotential2022, op.milky_way_potential_2022),
                      (gp.BovyMWPotential2014, op.bovy_mw_potential_2014)):
        s = cls().c_struct()
        ref = ofun()
        assert s.n == len(ref.components) and s.G == ref.G
        for i, c in enumerate(ref.components):
            assert s.c[i].kind == c.kind
            assert list(s.c[i].p)[: len(c.params)] == list(c.params)
    assert list(gp.MilkyWayPotential().keys()) == ["disk", "halo", "bulge", "nucleus"]
    assert list(gp.BovyMWPotential2014().keys()) == ["disk", "bulge", "halo"]
    custom = gp.MilkyWayPotential(disk=dict(m_tot=5e10), G=4.3e-12)
    assert custom["disk"].m_tot == 5e10 and custom.c_struct().G == 4.3e-12
    comp = gp.HernquistPotential(1e10, 1.0) + gp.NFWPotential(1e12, 20.0)
    assert comp.c_struct().n == 2


def test_unsupported_inputs_raise_not_fallback():
    import galax_b200.dynamics as gd
    import galax_b200.potential as gp

    with pytest.raises(NotImplementedError):
        gp.HernquistPotential(m_tot=lambda t: 1e10 * t, r_s=1.0).c_struct()
    pot = gp.MilkyWayPotential()
    with pytest.raises(NotImplementedError):
        gd.evaluate_orbit(pot, np.zeros((2, 6)), np.array([0.0, 1.0]), dense=True)
    with pytest.raises(TypeError):
        gd.evaluate_orbit("not a potential", np.zeros((2, 6)), np.array([0.0, 1.0]))


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry raises instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import galax_b200.dynamics as gd
    import galax_b200.potential as gp
    from galax_b200 import GalaxB200Error

    pot = gp.MilkyWayPotential()
    with pytest.raises(GalaxB200Error):
        pot.acceleration(np.ones((4, 3)))
    with pytest.raises(GalaxB200Error):
        gd.evaluate_orbit(pot, np.ones((4, 6)), np.array([0.0, 1.0]))


def test_product_code_never_imports_oracle():
    for f in (ROOT / "galax_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_xla_ffi_shim_type_checks_against_a_stub_of_the_ffi_header():
    """galax_b200/csrc/gx_xla_ffi.cc needs jaxlib's `xla/ffi/api/ffi.h`, which is not in this image.  tests/fake_xla holds
    a stand-in with the same public names whose binder checks, like the real one, that every handler's parameter list
    is exactly what its `Ffi::Bind()...` chain declares; the shim must compile against it (syntax + signatures), and a
    deliberately wrong binding must not."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    gxx = shutil.which("g++")
    cuda_inc = Path("/usr/local/cuda/include")
    if gxx is None or not (cuda_inc / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    src = root / "galax_b200" / "csrc" / "gx_xla_ffi.cc"
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-I", str(root / "tests" / "fake_xla"), "-I", str(cuda_inc)]
    ok = subprocess.run(cmd + [str(src)], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr
    bad = src.read_text().replace('.Attr<double>("t")', '.Attr<int32_t>("t")').replace(
        '"../../include/galax_b200.h"', f'"{root / "include" / "galax_b200.h"}"')
    res = subprocess.run(cmd + ["-x", "c++", "-"], input=bad, capture_output=True, text=True)
    assert res.returncode != 0 and "parameter list its binding declares" in res.stderr
