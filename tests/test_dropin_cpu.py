"""CPU: the `src` package is a drop-in for the import paths, class names and constructor signatures that the
reference's example scripts use.  The scripts themselves are read from /root/reference when present (build
container only; reference sources are never copied into this repository)."""
import ast
import importlib
import inspect
import os

import numpy as np
import pytest

from oracle import refshim

EXAMPLES = ['DamBreak', 'Containment', 'IceBreak']


def _tree(name):
    path = os.path.join(refshim.REFERENCE_ROOT, 'examples', name + '.py')
    return ast.parse(open(path).read())


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("example", EXAMPLES)
def test_example_imports_and_calls_bind(example):
    tree = _tree(example)
    names = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith('src'):
            mod = importlib.import_module(node.module)
            for a in node.names:
                assert hasattr(mod, a.name), (node.module, a.name)
                names[a.asname or a.name] = getattr(mod, a.name)
    assert 'Solver' in names and 'WCSPH' in names
    checked = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in names:
            obj = names[node.func.id]
            if inspect.isclass(obj):
                sig = inspect.signature(obj.__init__)
                sig.bind(None, *[None] * len(node.args), **{k.arg: None for k in node.keywords if k.arg})
                checked += 1
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and \
                isinstance(node.func.value, ast.Name) and node.func.value.id == 'Helpers':
            sig = inspect.signature(names['Helpers'].rect)
            sig.bind(*[None] * len(node.args), **{k.arg: None for k in node.keywords if k.arg})
            checked += 1
    assert checked >= 4


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_solver_attributes_used_by_icebreak_exist():
    """examples/IceBreak.py:111-171,271-284 read these from the solver / method objects."""
    from src.Solver import Solver
    from src.Methods.WCSPH import WCSPH
    from src.Integrators.PEC import PEC
    from src.Kernels.Wendland import Wendland
    s = Solver(WCSPH(1.0, 0.1, 1000.0, True, 0), PEC(True, False), Wendland(), 1.0)
    for a in ('couplingProperties', 't', 'settleTime', 'customSettle', 'dt', 'couplingIntegrator', 'kernel',
              'method', 'nn', 'export', 'particleArray', 'dt_a', 'dt_c', 'dt_f', 'timing_data', 'damping'):
        assert hasattr(s, a), a
    for a in ('rho0', 'gamma', 'B', 'r0', 'D', 'p1', 'p2', 'useSummationDensity', 'co', 'alpha', 'beta', 'epsilon'):
        assert hasattr(s.method, a), a
    assert callable(s.nn.nearPos) and callable(s.kernel.evaluate)


def test_helpers_rect_and_dtype():
    from src.Common import particle_dtype, computed_dtype, ParticleType, get_label_code
    from src.Helpers import Helpers
    assert particle_dtype.itemsize == 154 and computed_dtype.itemsize == 113
    assert [particle_dtype.fields[f][1] for f in ('deleted', 'label', 'm', 'rho0')] == [0, 1, 2, 146]
    assert get_label_code('temp-boundary') == ParticleType.TempBoundary
    p = Helpers.rect(0, 1, 0, 1, 0.25, pack=True)
    assert len(p) == 16 and p['y'][0] == 0.125 and p['y'][4] == 0.0
    q = Helpers.rect(0, 1, 0, 1, 0.25, pack=True, strict=True)
    assert len(q) == 14
    line = Helpers.rect(-1, -1, 0, 2, 0.5, label=ParticleType.Boundary)
    assert len(line) == 4 and np.all(line['label'] == 1)


def test_newmark_beta_host_integrator():
    from src.Integrators.NewmarkBeta import NewmarkBeta
    from src.Common import particle_dtype
    n = 3
    nb = NewmarkBeta(0.25, 0.5, np.eye(n), 2 * np.eye(n), np.zeros((n, n)))
    a = nb.acceleration(0.1, np.ones(n), np.zeros(n), np.zeros(n))
    assert np.allclose(a, 1.0 / (1 + 2 * 0.25 * 0.01))
    pA = np.zeros(n, dtype=particle_dtype); pA['label'] = 3; pA['vy'] = 1.0; pA['ay'] = 2.0
    pA = nb.predict(0.1, pA, 0.0)
    assert np.allclose(pA['y'], 0.1 + 2 * 0.25 * 0.01) and np.allclose(pA['vy'], 1.0 + 2 * 0.5 * 0.1)
