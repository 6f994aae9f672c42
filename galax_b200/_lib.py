"""Loader for ``libgalax_b200.so`` (the C ABI declared in ``include/galax_b200.h``).

There is no CPU fallback: if the library is missing or CUDA is unavailable, every compute entry
raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a (it cross-compiles on a
machine without a GPU).
"""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path
from typing import Sequence

_PKG = Path(__file__).resolve().parent
_SRC = _PKG / "csrc"
LIB_PATH = Path(os.environ.get("GALAX_B200_LIB", _PKG / "libgalax_b200.so"))  # override: A/B builds only
HEADER = _PKG.parent / "include" / "galax_b200.h"

NVCC_COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "20091",
    "-Xcompiler", "-fPIC",
]  # fmt: skip
# Development builds: GALAX_B200_FAST_BUILD=1 adds -split-compile 0 (2.3 min -> 1 min).  Not the default: without
# whole-module register allocation across the out-of-line right-hand side, C2's kernel measured 5 % slower.
if os.environ.get("GALAX_B200_FAST_BUILD") == "1":
    NVCC_COMPILE_FLAGS += ["-split-compile", "0"]

GX_MAX_COMPONENTS = 14
KIND_MN, KIND_HERNQUIST, KIND_NFW, KIND_PLC = 0, 1, 2, 3
KIND_LOG, KIND_ISOCHRONE, KIND_SATOH = 4, 5, 6
KIND_TRIAXIAL_HERNQUIST, KIND_JAFFE, KIND_BURKERT, KIND_STONE, KIND_HARMONIC, KIND_HENON_HEILES = 7, 8, 9, 10, 11, 12
PHI, GRAD, ACC, HESS = 1, 2, 4, 8
OK, MAX_STEPS_REACHED, NONFINITE = 0, 1, 2
SCHEME_SIE, SCHEME_LEAPFROG_MIDPOINT = 0, 1
SCHEME_GENERAL_KERNEL = 0x100
SCHEME_STRICT = 0x200
LAYOUT_NT3, LAYOUT_T3N = 0, 1
DF_FARDAL15, DF_CHEN24 = 0, 1
DENSE_RECORD_DOUBLES = 51
SOLVER_DOPRI8, SOLVER_DOPRI5 = 8, 5
SOLVER_STRICT = 0x200


class GxComponent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 8), ("dp", C.c_double * 8)]


class GxPotential(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved", C.c_int32), ("G", C.c_double), ("c", GxComponent * GX_MAX_COMPONENTS)]


class GxPid(C.Structure):
    _fields_ = [
        ("rtol", C.c_double), ("atol", C.c_double),
        ("pcoeff", C.c_double), ("icoeff", C.c_double), ("dcoeff", C.c_double),
        ("safety", C.c_double), ("factormin", C.c_double), ("factormax", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double),
        ("force_dtmin", C.c_int32), ("reserved", C.c_int32),
        ("dt0", C.c_double),
    ]  # fmt: skip


class GxOrbitEpilogue(C.Structure):
    _fields_ = [("energy", C.c_void_p), ("angmom", C.c_void_p), ("tidal", C.c_void_p)]


class GalaxB200Error(RuntimeError):
    pass


def sources() -> list[tuple[Path, list[str]]]:
    """Translation units and their extra flags.  ``gx_strict.cu`` (the reference-order kernels) is compiled without
    floating-point contraction: its arithmetic must be reproducible bit for bit by a plain C program."""
    return [(_SRC / "gx_kernels.cu", []), (_SRC / "gx_strict.cu", ["-fmad=false"])]


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    m = LIB_PATH.stat().st_mtime
    deps = list(_SRC.glob("*.cu")) + list(_SRC.glob("*.cuh")) + list(_SRC.glob("*.h")) + list(HEADER.parent.glob("*.h"))
    return any(d.stat().st_mtime > m for d in deps)


def build(force: bool = False, verbose: bool = False, defines: Sequence[str] = (), out: Path | None = None) -> Path:
    """Compile ``libgalax_b200.so`` for sm_100a if it is missing or older than its sources."""
    target = Path(out) if out is not None else LIB_PATH
    if out is None and not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise GalaxB200Error("nvcc not found: cannot build libgalax_b200.so")
    env = dict(os.environ)
    env.pop("CC", None)  # the image's $CC points at a gcc nvcc cannot drive
    env.pop("CXX", None)
    objdir = _PKG / "build"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for src, extra in sources():
        obj = objdir / (target.stem + "_" + src.stem + ".o")
        cmd = [nvcc, *NVCC_COMPILE_FLAGS, *extra, *defines, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        jobs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)))
    objs = []
    for obj, proc in jobs:
        out_, err_ = proc.communicate()
        if proc.returncode != 0:
            raise GalaxB200Error(f"nvcc failed:\n{out_}\n{err_}")
        if verbose:
            print(err_)
        objs.append(str(obj))
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(target), *objs],
                         capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise GalaxB200Error(f"nvcc link failed:\n{res.stdout}\n{res.stderr}")
    for o in objs:  # (the objects are as large as the library and would travel with every snapshot of the tree)
        Path(o).unlink(missing_ok=True)
    return target


_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "gx_version": (C.c_int, []),
    "gx_strerror": (C.c_char_p, [C.c_int]),
    "gx_workspace_bytes": (C.c_int64, []),
    "gx_potential_eval": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_double, C.c_int64, C.c_uint32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_integrate_fixed": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                     C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_integrate_fixed_epilogue": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                              C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GxOrbitEpilogue),
                                              C.c_void_p]),
    "gx_integrate_adaptive_epilogue": (C.c_int, [C.c_int32, C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p,
                                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                                 C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.POINTER(GxOrbitEpilogue), C.c_void_p]),
    "gx_joint_workspace_bytes": (C.c_int64, [C.c_int64]),
    "gx_integrate_adaptive_joint": (C.c_int, [C.c_int32, C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p, C.c_void_p,
                                              C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int64,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "gx_integrate_dopri8": (C.c_int, [C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int64,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_integrate_dopri8_record": (C.c_int, [C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p, C.c_void_p,
                                             C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_dense_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "gx_integrate_adaptive": (C.c_int, [C.c_int32, C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int64,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_integrate_adaptive_record": (C.c_int, [C.c_int32, C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p,
                                               C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_dense_eval_solver": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_stream_release": (C.c_int, [C.POINTER(GxPotential), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "gx_stream_release_t": (C.c_int, [C.POINTER(GxPotential), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_energy_angmom_t": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_energy_angmom": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "gx_host_potential_eval": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_double, C.c_int64, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_host_integrate_fixed": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                          C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_host_integrate_dopri8": (C.c_int, [C.POINTER(GxPotential), C.POINTER(GxPid), C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int32,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_bench_dfma": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "gx_jax_normal": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]),
    "gx_jax_fardal_chain": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gx_jax_fardal_chain_workspace_bytes": (C.c_int64, [C.c_int64]),
    "gx_debug_math": (C.c_int, [C.c_int32, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "gx_force_table": (C.c_int, [C.c_int32, C.c_double, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "gx_spherical_force_table": (C.c_int, [C.POINTER(GxPotential), C.c_void_p, C.c_int64, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                           C.POINTER(C.c_double)]),
    "gx_fixed_time_grid": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int32]),
}  # fmt: skip

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """The loaded shared library (loaded once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise GalaxB200Error(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for this path)"
            )
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().gx_strerror(rc).decode()
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise GalaxB200Error(f"{what}: {msg} (code {rc})")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise GalaxB200Error("galax_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
