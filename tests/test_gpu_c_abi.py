"""The C ABI used from plain C (examples/c_abi_demo.c): compiled with gcc against include/galax_b200.h, linked with
libgalax_b200.so, no CUDA headers and no Python in the call path; its printed results must equal the Python mirror's."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _build(tmp_path):
    exe = tmp_path / "c_abi_demo"
    lib_dir = ROOT / "galax_b200"
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "c_abi_demo.c"),
           f"-L{lib_dir}", "-lgalax_b200", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", str(exe)]  # fmt: skip
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_plain_c_caller_compiles_and_links(built_lib, tmp_path):
    """CPU: the header is valid C99 and every symbol the demo uses resolves (no run: needs a GPU)."""
    assert _build(tmp_path).exists()


@pytest.mark.gpu
def test_plain_c_caller_matches_python_mirror(built_lib, tmp_path):
    import galax_b200.dynamics as gd
    import galax_b200.potential as gp

    out = subprocess.run([str(_build(tmp_path))], check=True, capture_output=True, text=True).stdout.splitlines()
    rows = {ln.split()[0] + (ln.split()[1] if ln.split()[0] in ("sie", "dopri8") else ""): ln.split() for ln in out}
    pot = gp.MilkyWayPotential()
    x = np.array([1.0, 2.0, 3.0])
    assert np.isclose(float(rows["phi"][1]), pot.potential(x), rtol=1e-12)
    assert np.allclose([float(v) for v in rows["grad"][1:]], pot.gradient(x), rtol=1e-12)
    assert np.allclose([float(v) for v in rows["hess_diag"][1:]], np.diag(pot.hessian(x)), rtol=1e-11)
    i = np.arange(16)
    r, ang = 4.0 + i, 0.37 * i
    q0 = np.stack([r * np.cos(ang), r * np.sin(ang), 0.1 * i - 0.5], axis=1)
    p0 = np.stack([-0.2 * np.sin(ang), 0.2 * np.cos(ang), 0.01 * i], axis=1)
    sie = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
    qs = sie.solve(pot, (q0, p0), 0.0, 1000.0, dt0=0.1).ys[0][:, 0]
    qd = gd.OrbitSolver().solve(pot, (q0, p0), 0.0, 1000.0).ys[0][:, 0]
    for k in (0, 5, 10, 15):
        assert int(rows[f"sie{k}"][2]) == 0 and int(rows[f"dopri8{k}"][2]) == 0
        assert np.allclose([float(v) for v in rows[f"sie{k}"][3:]], qs[k], rtol=1e-11, atol=1e-11)
        assert np.allclose([float(v) for v in rows[f"dopri8{k}"][5:]], qd[k], rtol=1e-9, atol=1e-9)
