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
def test_threaded_loop_is_the_same_loop(name):
    """oracle_loop_mt (bench.py's CPU legs: the pair loop on all host threads) against oracle_loop: every particle's
    arithmetic and its order are the same, so the results must be bit-identical -- also with a strided sample and with
    more threads than blocks of particles."""
    g, meta, pA = load_golden(name)
    for stride, threads in ((1, 3), (1, 64), (4, 5)):
        P1, Pn = O.Particles.from_aos(pA), O.Particles.from_aos(pA)
        g1, gn = O.Grid(P1, meta['scale']), O.Grid(Pn, meta['scale'])
        pairs1 = O.loop(P1, _wcsph(meta), g1, meta['kernel'], stride, 1 % stride)
        pairsn = O.loop(Pn, _wcsph(meta), gn, meta['kernel'], stride, 1 % stride, threads=threads)
        assert pairs1 == pairsn
        for f in ('rho', 'p', 'c', 'drho', 'ax', 'ay', 'xsphx', 'xsphy'):
            assert np.array_equal(getattr(P1, f), getattr(Pn, f), equal_nan=True), (f, stride, threads)


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


# ---- the equations one by one, on the tables of the reference's own equation tests --------------------------------
def _leaf():
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if here not in sys.path:
        sys.path.insert(0, here)
    import leaf_inputs as LI
    return LI, np.load(os.path.join(here, "equations_leaf.npz"))


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_momentum_leaf_vs_reference_and_its_own_test():
    """reference test/test_numba_momentum.py:21-90: Momentum(0.01, 0, pa, comp) on 10 000 linspace neighbours equals the
    test's vectorised closed form (_calc_old, :92-123); golden = the reference function itself on the same table."""
    LI, g = _leaf()
    pa, comp = LI.momentum_reftest()
    a = O.eq_momentum(0.01, 0.0, pa, comp)
    assert _rel(a, g['momentum_reftest']) <= 1e-13
    # the closed form of the reference test: pressure part only here (every v.x >= 0, no viscosity)
    tmp = pa['p'] / pa['rho'] ** 2 + comp['p'] / comp['rho'] ** 2
    assert a[0] == pytest.approx(np.sum(-comp['m'] * tmp * comp['dw_x']), rel=1e-12)
    assert a[1] == pytest.approx(np.sum(-comp['m'] * tmp * comp['dw_y']), rel=1e-12)
    # mixed labels, approaching pairs (artificial viscosity with alpha, and with beta != 0)
    pa, comp = LI.random_table()
    assert _rel(O.eq_momentum(0.01, 0.0, pa, comp), g['random_momentum']) <= 1e-13
    assert _rel(O.eq_momentum(0.3, 0.7, pa, comp), g['random_momentum_beta']) <= 1e-13
    fl = comp['label'] == 0
    dot = comp['vx'] * comp['x'] + comp['vy'] * comp['y']
    hij = 0.5 * (pa['h'] + comp['h']); cij = 0.5 * (pa['c'] + comp['c']); rij = 0.5 * (pa['rho'] + comp['rho'])
    mu = hij * dot / (comp['r'] ** 2 + 0.01 * hij ** 2)
    pi = np.where(dot < 0, mu * (0.7 * mu - 0.3 * cij) / rij, 0.0)
    fac = pa['p'] / pa['rho'] ** 2 + comp['p'] / comp['rho'] ** 2 + pi
    assert O.eq_momentum(0.3, 0.7, pa, comp)[0] == pytest.approx(np.sum((-comp['m'] * fac * comp['dw_x'])[fl]), rel=1e-12)


def test_continuity_leaf_vs_reference_and_its_own_test():
    """reference test/test_numba_continuity.py:12-66 (old_calc: sum m (v . dW)); 100 000 neighbours instead of 10 M."""
    LI, g = _leaf()
    comp = LI.continuity_reftest()
    d = O.eq_continuity(comp)
    assert abs(d / float(g['continuity_reftest']) - 1) <= 1e-12
    assert abs(d / np.sum(comp['m'] * (comp['vx'] * comp['dw_x'] + comp['vy'] * comp['dw_y'])) - 1) <= 1e-12
    _, comp = LI.random_table()
    assert abs(O.eq_continuity(comp) / float(g['random_continuity']) - 1) <= 1e-13


def test_xsph_and_boundary_force_leaves_vs_reference():
    """reference src/Equations/XSPH.py:6-31 and BoundaryForce.py:7-42 (test/test_eq_boundary.py:9-26: the wall on the left
    pushes to +x, nothing in y), also with the general exponents and a coincident wall particle (r <= 1e-12 skipped)."""
    LI, g = _leaf()
    f = O.eq_boundary_force(1.0, 5 * 9.81 * 1.0, 4.0, 2.0, LI.boundary_reftest())
    assert f[0] > 0 and f[1] == 0.0
    assert _rel(f, g['boundary_reftest']) <= 1e-14
    pa, comp = LI.random_table()
    assert _rel(O.eq_xsph(0.5, pa, comp), g['random_xsph']) <= 1e-13
    assert _rel(O.eq_boundary_force(0.15, 1226.25, 4.0, 2.0, comp), g['random_boundary_42']) <= 1e-13
    assert _rel(O.eq_boundary_force(0.15, 1226.25, 12.0, 6.0, comp), g['random_boundary_126']) <= 1e-13


def test_courant_known_answers():
    """reference test/test_eq_courant.py:8-11 (0, 0.4, 0.2) and the reference function on a random table."""
    LI, g = _leaf()
    assert O.eq_courant(0.4, [0.0], [1.0]) == 0.0
    assert O.eq_courant(0.4, [1.0], [1.0]) == pytest.approx(0.4, abs=1e-15)
    assert O.eq_courant(0.4, [2.0], [4.0]) == pytest.approx(0.2, abs=1e-15)
    _, comp = LI.random_table()
    got = [O.eq_courant(0.4, [0.0], [1.0]), O.eq_courant(0.4, [1.0], [1.0]), O.eq_courant(0.4, [2.0], [4.0]),
           O.eq_courant(0.25, comp['h'], comp['c'])]
    assert np.allclose(got, g['courant'], rtol=1e-15, atol=0)


def test_gaussian_matches_the_reference_tests_old_implementation():
    """reference test/test_numba_kernel.py:13-71: Gaussian.evaluate against the test's numpy `old_func`
    (alpha / h^2 exp(-q^2) for q <= 3, else 0) on a diagonal line of 1 000 points with h = 1.3 r."""
    x = np.linspace(0, 1000, 1000)
    pts = np.stack([x, x], axis=1)
    for i in (0, 1, 500, 999):
        d = np.delete(np.hypot(*(pts[i] - pts).T), i)
        h = d * 1.3
        q = d / h
        old = np.where(q <= 3, (1 / np.pi) / h ** 2 * np.exp(-q * q), 0.0)
        assert np.allclose(O.kernel_evaluate('gaussian', d, h), old, rtol=1e-14, atol=0)


def test_leaves_compose_to_the_loop():
    """The stand-alone leaves, fed from the oracle's own neighbour query and kernels, reproduce oracle_loop (the composed
    step that the golden vectors pin) particle by particle: the per-pair formulas pinned above ARE the ones in the loop."""
    g, meta, pA = load_golden('dambreak20_wendland')
    LI, _ = _leaf()
    P = O.Particles.from_aos(pA)
    w = _wcsph(meta)
    grid = O.Grid(P, 2.0)
    Q = P.copy()
    O.loop(Q, w, grid, meta['kernel'])
    fl = np.flatnonzero(P.label == 0)
    for i in list(fl[:40]) + list(fl[-40:]):
        hij, qij, r, idx = grid.near(int(i))
        comp = np.zeros(len(idx), dtype=LI.COMP_DTYPE)
        comp['label'] = P.label[idx]
        for f_ in ('m', 'rho', 'p'):
            comp[f_] = getattr(Q, f_)[idx]               # p after the EOS pass of the loop
        comp['h'] = hij; comp['r'] = r; comp['q'] = qij     # comp.c stays 0 (SolverTools.py:90)
        comp['x'] = P.x[i] - P.x[idx]; comp['y'] = P.y[i] - P.y[idx]
        comp['vx'] = P.vx[i] - P.vx[idx]; comp['vy'] = P.vy[i] - P.vy[idx]
        comp['w'] = O.kernel_evaluate(meta['kernel'], r, hij)
        comp['dw_x'] = O.kernel_gradient(meta['kernel'], comp['x'], r, hij)
        comp['dw_y'] = O.kernel_gradient(meta['kernel'], comp['y'], r, hij)
        self_ = dict(p=Q.p[i], rho=P.rho[i], h=P.h[i], c=Q.c[i])
        a = O.eq_momentum(w.alpha, w.beta, self_, comp)
        b = O.eq_boundary_force(w.r0, w.D, w.p1, w.p2, comp)
        xs = O.eq_xsph(w.epsilon, self_, comp)
        scale = max(abs(Q.ax[i]), abs(Q.ay[i]), 1.0)
        assert abs(a[0] + b[0] - Q.ax[i]) <= 1e-12 * scale and abs(a[1] - 9.81 + b[1] - Q.ay[i]) <= 1e-12 * scale
        assert abs(O.eq_continuity(comp) - Q.drho[i]) <= 1e-12 * max(abs(Q.drho[i]), 1.0)
        assert abs(P.vx[i] + xs[0] - Q.xsphx[i]) <= 1e-13 * max(abs(Q.xsphx[i]), 1.0)
        assert abs(P.vy[i] + xs[1] - Q.xsphy[i]) <= 1e-13 * max(abs(Q.xsphy[i]), 1.0)
