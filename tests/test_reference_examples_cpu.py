"""CPU (build container only): the reference's OWN example scripts -- examples/DamBreak.py, Containment.py and
IceBreak.py, byte for byte as shipped -- run against this package.

The scripts are copied at test time from /root/reference into a scratch directory laid out as INTEGRATION.md section 1
describes (this package's `src/` and `osph_b200/` next to `examples/`), so that their unedited
`sys.path.insert(1, <examples>/..)` + `from src.Solver import Solver` resolve to the re-hosted API; nothing of the
reference is stored in this repository.  The device library is the SIMT-emulated build of the kernel sources
(tests/emu), and OSPH_MAX_STEPS cuts the runs to a dozen steps.  What is checked: the scripts construct WCSPH / PEC /
Verlet / NewmarkBeta / kernels / Solver with their own arguments, settle, step (IceBreak through its coupling callback
with nearPos, Shepard and summation density on the coupled ice row) and save, without edits.
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import refshim

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))
import build as emu_build  # noqa: E402

pytestmark = pytest.mark.skipif(not (refshim.available() and emu_build.available()),
                                reason="needs the reference tree and g++ / CUDA headers for the emulated library")
STEPS = 12


@pytest.fixture(scope="module")
def layout(tmp_path_factory):
    top = tmp_path_factory.mktemp("dropin")
    shutil.copytree(os.path.join(refshim.REFERENCE_ROOT, "examples"), top / "examples",
                    ignore=shutil.ignore_patterns("*.hdf5", "*.mp4", "__pycache__"))
    os.symlink(os.path.join(ROOT, "offshore-sph_b200", "src"), top / "src")
    os.symlink(os.path.join(ROOT, "offshore-sph_b200", "osph_b200"), top / "osph_b200")
    stubs = top / "stubs"
    stubs.mkdir()
    # the image has no h5py; IceBreak.py imports it at module level (and uses it only in its own post-processing)
    (stubs / "h5py.py").write_text("# namespace stand-in: h5py is not installed in this image\n")
    # numba < 0.59 ran `@numba.jit()` helpers that take Python objects in object mode; IceBreak.py relies on that
    (stubs / "sitecustomize.py").write_text(
        "import numba\n"
        "def _jit(*a, **k):\n"
        "    if a and callable(a[0]):\n"
        "        return a[0]\n"
        "    return lambda f: f\n"
        "numba.jit = _jit\n")
    env = dict(os.environ, OSPH_LIB=emu_build.build(), OSPH_MAX_STEPS=str(STEPS), OSPH_QUIET="1",
               PYTHONPATH=str(stubs), PYTHONDONTWRITEBYTECODE="1")
    return top, env


def _run(layout, script, stdin=""):
    top, env = layout
    return subprocess.run([sys.executable, os.path.join("examples", script)], cwd=top, env=env, input=stdin,
                          capture_output=True, text=True, timeout=600)


def _saved(path):
    z = np.load(path)
    from oracle.oracle import particle_dtype
    pA = np.frombuffer(z['particleArray'].tobytes(), dtype=particle_dtype)
    return z, pA


def test_dam_break_script_runs_unedited(layout):
    out = _run(layout, "DamBreak.py")
    assert out.returncode == 0, out.stderr[-2000:]
    z, pA = _saved(layout[0] / "examples" / "dam-break-2d.hdf5.npz")
    assert len(z['dt_a']) == STEPS and np.all(z['dt_a'] > 0)
    live = ~pA['deleted']
    fluid = live & (pA['label'] == 0)
    assert fluid.sum() == 2500 and live.sum() == 2983                 # N = 50 as shipped
    for f in ('x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay'):
        assert np.all(np.isfinite(pA[f][live])), f
    assert np.all(pA['rho'][fluid] > 900.0) and np.any(pA['ay'][fluid] != 0.0)


def test_containment_script_runs_unedited(layout):
    """Ships with plot = True and calls WCSPH without useSummationDensity: both must go through."""
    out = _run(layout, "Containment.py")
    assert out.returncode == 0, out.stderr[-2000:]
    z, pA = _saved(layout[0] / "examples" / "containment.hdf5.npz")
    assert len(z['dt_a']) == STEPS
    live = ~pA['deleted']
    assert np.all(np.isfinite(pA['x'][live])) and np.all(pA['h'][live & (pA['label'] == 0)] > 0)    # dynamic h


def test_ice_break_script_runs_unedited_up_to_its_own_post_processing(layout):
    """Model 1 (short floating ice sheet), no animation, 30 particles across: settling, the coupled ice row advanced by
    NewmarkBeta inside the script's coupling callback (pressure probe via solver.nn.nearPos + Shepard), save.  The
    script then re-opens its output with h5py itself, which this image does not have: that is where it stops."""
    out = _run(layout, "IceBreak.py", stdin="1\nn\n30\n")
    saved = layout[0] / "examples" / "IceBreak-1.hdf5.npz"
    assert saved.exists(), out.stderr[-2000:]
    assert out.returncode == 0 or ("in post" in out.stderr and "h5py" in out.stderr), out.stderr[-2000:]
    z, pA = _saved(saved)
    assert len(z['dt_a']) == STEPS
    live = ~pA['deleted']
    assert (pA['label'][live] == 3).sum() > 0                              # the Coupled ice row is part of the run
    for f in ('x', 'y', 'vx', 'vy', 'rho'):
        assert np.all(np.isfinite(pA[f][live])), f
