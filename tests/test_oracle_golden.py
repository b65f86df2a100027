"""CPU: pin the oracle (oracle/wcsph_oracle.c) against golden vectors frozen from the reference.

Bit-exact: grid parameters, cell ids, neighbour sets AND order, dt triples of the first step.
Field math (fastmath numba in the reference vs strict IEEE here): <= 1e-12 norm-wise.
"""
import numpy as np
import pytest

from conftest import STEP_CASES, field_err, load_golden
from oracle import oracle as O

TOL = 1e-12


def _wcsph(meta):
    c = meta['consts']
    w = O.wcsph(c['height'], c['r0'], c['rho0'], meta['useXSPH'], c['Pb'], meta.get('summation', False))
    assert w.co == c['co'] and w.B == c['B'] and w.D == c['D']
    return w


@pytest.mark.parametrize("name", STEP_CASES)
def test_grid_cells_neighbours_bit_exact(name):
    g, meta, pA = load_golden(name)
    P = O.Particles.from_aos(pA)
    grid = O.Grid(P, meta['scale'])
    p = grid.params
    assert grid.rc == 0
    assert [p['xmin'], p['xmax'], p['ymin'], p['ymax'], p['cell_size']] == list(g['grid'][:5])
    assert (p['ncx'], p['ncy']) == (int(g['grid'][5]), int(g['grid'][6]))
    assert np.array_equal(grid.cell_ids(), g['cell_ids'])
    off, idx = grid.neighbours_csr()
    assert np.array_equal(off, g['nbr_off'])
    assert np.array_equal(idx, g['nbr_idx'].astype(np.int64))      # same order, not only same set


@pytest.mark.parametrize("name", STEP_CASES)
def test_loop_fields(name):
    g, meta, pA = load_golden(name)
    P = O.Particles.from_aos(pA)
    grid = O.Grid(P, meta['scale'])
    pairs = O.loop(P, _wcsph(meta), grid, meta['kernel'])
    assert pairs == len(g['nbr_idx'])
    for f in ('p', 'c', 'drho', 'ax', 'ay', 'xsphx', 'xsphy'):
        assert field_err(getattr(P, f), g['loop_' + f]) <= TOL, f


@pytest.mark.parametrize("name", STEP_CASES)
def test_whole_steps(name):
    g, meta, pA = load_golden(name)
    P = O.Particles.from_aos(pA)
    w = _wcsph(meta)
    for s in range(meta['nsteps']):
        dt3, _ = O.step(P, w, meta['kernel'], 'pec', meta['useXSPH'], meta['strict'], meta['damping'],
                        meta['fixed_h'])
        if s == 0:
            assert list(dt3) == list(g['dts'][0])               # strict-IEEE reduction of identical inputs
        else:
            assert np.allclose(dt3, g['dts'][s], rtol=1e-11, atol=0)
        for f in ('x', 'y', 'vx', 'vy', 'rho', 'drho', 'ax', 'ay', 'xsphx', 'xsphy', 'p', 'h'):
            assert field_err(getattr(P, f), g['step_' + f][s]) <= 1e-11, (s, f)


# ---- known answers of the reference's own unit tests --------------------------------------------

def test_cubic_known_answers():
    """reference test/test_kernels_cubic.py:11-72 (closed forms at q = 0.5, 1, 1.5, 2, 3)."""
    a = 10 / (7 * np.pi)
    r = np.array([0.5, 1.0, 1.5, 2.0, 3.0]); h = np.ones(5)
    w = O.kernel_evaluate('cubic', r, h)
    assert np.allclose(w, [a * (1 - 1.5 * .25 * (1 - .25)), a * .25, a * .25 * .5 ** 3, 0, 0], rtol=1e-15)
    g = O.kernel_gradient('cubic', r, r, h)     # x = r: derivative along the separation
    assert np.allclose(g, [a * -3 * .5 * (1 - .375), a * -.75, a * -.75 * .25, 0, 0], rtol=1e-15)
    assert O.kernel_gradient('cubic', np.array([1e-11]), np.array([1e-11]), np.ones(1))[0] == 0.0


def test_gaussian_known_answers():
    """reference test/test_kernels_Gaussian.py:28-101 / test/PySPH/Gaussian.py (2-D Gaussian, cut at q=3)."""
    rng = np.random.default_rng(0)
    for h in (0.5, 1.0, 1.7):
        r = rng.uniform(0, 4 * h, 200); x = rng.uniform(-1, 1, 200) * r
        hh = np.full(200, h); q = r / h
        w = np.where(q <= 3, np.exp(-q * q) / (np.pi * h * h), 0)
        assert np.allclose(O.kernel_evaluate('gaussian', r, hh), w, rtol=1e-14)
        dw = np.where(q <= 3, -2 * q * np.exp(-q * q) / (np.pi * h * h) / (r * h) * x, 0)
        assert np.allclose(O.kernel_gradient('gaussian', x, r, hh), dw, rtol=1e-13)


def test_wendland_normalisation():
    """Wendland has no reference test; check the 2-D volume integral is 1 and the gradient is dW/dr."""
    h = 0.7
    r = np.linspace(0, 2 * h, 200001)
    w = O.kernel_evaluate('wendland', r, np.full_like(r, h))
    assert abs(np.trapezoid(2 * np.pi * r * w, r) - 1.0) < 1e-8
    dw = O.kernel_gradient('wendland', r, r, np.full_like(r, h))[1:-1]
    num = (w[2:] - w[:-2]) / (r[2:] - r[:-2])
    assert np.max(np.abs(dw - num)) < 1e-6 * np.max(np.abs(dw))


def test_pec_known_answers():
    """reference test/test_integrators_pec.py:13-45 (expected values; the test file itself is stale)."""
    for label, moved in ((1, False), (0, True)):
        P = O.Particles(1)
        P.label[0] = label; P.vx[0] = 1.0; P.vy[0] = 3.0; P.ax[0] = 5.0; P.drho[0] = 10.0
        dt = 2.0
        mask = np.ones(1, dtype=np.uint8) if moved else np.zeros(1, dtype=np.uint8)
        O.pec_predict(P, mask, dt, 0.0, useXSPH=False)
        if moved:
            assert (P.x[0], P.y[0], P.rho[0]) == (dt * 0.5, dt * 0.5 * 3.0, dt * 0.5 * 10.0)
        else:   # the Solver never hands boundary rows to the integrator (src/Solver.py:380)
            assert (P.x[0], P.y[0]) == (0.0, 0.0)
        O.pec_correct(P, mask, dt, 0.0, useXSPH=False)
        if moved:
            assert P.x[0] == pytest.approx(dt * 1.0 + 0.5 * 5 * 2 * 2)
            assert P.y[0] == pytest.approx(dt * 3.0)
            assert P.rho[0] == pytest.approx(10 * dt)


def test_pec_equals_euler_with_frozen_forces():
    """reference test/test_integrators_pec.py:47-65."""
    A = O.Particles(1); A.vx[0] = 1.0; A.vy[0] = 3.0; A.ax[0] = 5.0; A.drho[0] = 10.0
    B = A.copy()
    m = np.ones(1, dtype=np.uint8)
    O.pec_predict(A, m, 2.0, 0.0, useXSPH=False); O.pec_correct(A, m, 2.0, 0.0, useXSPH=False)
    O.euler_correct(B, m, 2.0)
    for f in ('x', 'y', 'vx', 'vy', 'rho'):
        assert getattr(A, f)[0] == pytest.approx(getattr(B, f)[0])


def test_euler_known_answer():
    """reference test/test_integrators_euler.py:30-38 (the fluid half; the class itself no longer builds under numba
    >= 0.59, the boundary half of that test contradicts Euler.py:17-26 and is stale)."""
    P = O.Particles(1)
    P.vx[0] = 1.0; P.vy[0] = 3.0; P.ax[0] = 5.0; P.drho[0] = 10.0
    dt = 2.0
    O.euler_correct(P, np.ones(1, dtype=np.uint8), dt)
    assert P.x[0] == 1.0 * dt + 0.5 * 5.0 * dt * dt and P.y[0] == 3.0 * dt and P.rho[0] == 10 * dt
    assert (P.vx[0], P.vy[0]) == (1.0 + dt * 5.0, 3.0)


def test_verlet_matches_reference_verlet():
    """Golden `verlet_leaf` = the reference's own Verlet.predict / .correct (src/Integrators/Verlet.py:28-55, run
    under numba by tests/golden/make_golden.py), with and without XSPH.  No density update, no damping: as shipped."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'verlet_leaf.npz'))
    pA = np.frombuffer(g['aos'].tobytes(), dtype=O.particle_dtype)
    dt = float(g['dt'])
    mask = np.ones(len(pA), dtype=np.uint8)
    for tag, xs in (('xsph', True), ('raw', False)):
        P = O.Particles.from_aos(pA)
        O.verlet_predict(P, mask, dt)
        want = np.frombuffer(g['pred_' + tag].tobytes(), dtype=O.particle_dtype)
        for f in ('x', 'y', 'vx', 'vy', 'rho'):
            assert np.array_equal(getattr(P, f), want[f]), (tag, 'predict', f)
        O.verlet_correct(P, mask, dt, useXSPH=xs)
        want = np.frombuffer(g['corr_' + tag].tobytes(), dtype=O.particle_dtype)
        for f in ('x', 'y', 'vx', 'vy', 'rho'):
            assert field_err(getattr(P, f), want[f]) <= 1e-15, (tag, 'correct', f)
        assert np.array_equal(P.rho, pA['rho'])


def test_linked_list_known_answer():
    """reference test/test_linked_list.py:56-79: 26x26 unit lattice, h = 1, scale = 3."""
    xv = np.linspace(0, 25, 26)
    x, y = np.meshgrid(xv, xv, indexing='ij')
    P = O.Particles(26 * 26)
    P.x[:] = x.ravel(); P.y[:] = y.ravel(); P.h[:] = 1.0
    grid = O.Grid(P, 3.0)
    assert grid.params['cell_size'] == 3.0
    near0 = set(grid.near(0)[3].tolist())
    assert {0, 1, 26, 27} <= near0
    q = np.hypot(P.x - P.x[0], P.y - P.y[0]) / 1.0
    assert near0 == set(np.flatnonzero(q <= 3.0).tolist())


def test_boundary_force_direction():
    """reference test/test_eq_boundary.py:9-26: a wall particle on the left pushes the fluid to +x."""
    P = O.Particles(3)
    P.label[:] = [0, 1, 0]
    P.x[:] = [0.0, -1.0, 1.0]; P.h[:] = [1.0, 0.0, 1.0]
    P.m[:] = [0.0, 0.0, 0.0]; P.rho[:] = [1000.0, 0.0, 1000.0]
    w = O.wcsph(1.0, 2.0, 1000.0, False)
    grid = O.Grid(P, 2.0)
    O.loop(P, w, grid, 'cubic')
    assert P.ax[0] > 0 and P.ay[0] == -9.81


def test_kinetic_energy_and_timestep():
    """reference test/test_kinetic_energy.py:11-23 (100 unit particles -> 50) and TimeStep.py:40-56."""
    P = O.Particles(100)
    P.m[:] = 1.0; P.vx[:] = 1.0
    assert O.kinetic_energy(P) == 50.0
    P.h[:] = 0.8; P.c[:] = 200.0
    assert O.timestep(P, P.fluid) == (0.25 * 0.8 / 200.0, 0.25 * 0.8 / 200.0, 1e10)
    P.ax[3] = 3.0; P.ay[3] = 4.0                     # the force criterion uses a^2, not |a|
    assert O.timestep(P, P.fluid)[2] == 0.25 * np.sqrt(0.8 / 25.0)


def test_tait():
    """reference test/test_numba_taiteos.py:41-46 (formula) and WCSPH constants."""
    w = O.wcsph(25.0, 0.5, 1000.0, True)
    assert w.co == 221.47234590350104 and w.B == 7007142.857142859 and w.D == 1226.25
    L = O.lib()
    assert L.oracle_tait_p(7.0, w.B, 1000.0, 1000.0, 0) == 0.0
    assert L.oracle_tait_p(7.0, w.B, 1000.0, 1010.0, 1) == 0.0
    assert L.oracle_tait_p(7.0, w.B, 1000.0, 1010.0, 0) == pytest.approx(w.B * (1.01 ** 7 - 1), rel=1e-14)
