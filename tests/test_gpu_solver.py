"""GPU: the re-hosted Solver and the plugin objects used stand-alone, against the reference's own
Solver.run() (golden end state) and the oracle."""
import numpy as np
import pytest

from conftest import field_err, load_golden
from oracle import oracle as O
from osph_b200 import workloads as W

pytestmark = pytest.mark.gpu


def test_solver_run_matches_reference_solver(monkeypatch):
    """Settling -> gate removal -> time stepping of the 12x12 dam break, vs the unmodified reference Solver."""
    monkeypatch.setenv("OSPH_QUIET", "1")
    from src.Solver import Solver
    from src.Methods.WCSPH import WCSPH
    from src.Kernels.Wendland import Wendland
    from src.Integrators.PEC import PEC
    g, meta, _ = load_golden('solver_dambreak12_wendland')
    r0, pA = W.dam_break(meta['N'])
    method = WCSPH(height=25.0, r0=r0, rho0=1000.0, useXSPH=True, Pb=0, useSummationDensity=False)
    s = Solver(method, PEC(useXSPH=True, strict=False), Wendland(), meta['duration'], incrementalWriteout=False,
               h=meta['hfac'] * r0, maxSettle=meta['maxSettle'])
    s.addParticles(pA)
    s.setup()
    s.run()
    ref = np.frombuffer(g['final'].tobytes(), dtype=O.particle_dtype)
    assert s.t_step == int(g['t_step'])
    assert np.allclose(s.dt_a, g['dt_a'], rtol=1e-10, atol=0)
    assert s.settleTime == pytest.approx(float(g['settleTime']), rel=1e-10)
    assert s.t == pytest.approx(float(g['t']), rel=1e-10)
    assert np.array_equal(s.particleArray['deleted'], ref['deleted'])
    act = ~ref['deleted']
    for f in ('x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay', 'drho', 'h', 'c'):
        assert field_err(s.particleArray[f][act], ref[f][act]) <= 1e-9, f
    assert np.all(s.particleArray['p'][~act] == -1e15)
    assert len(s.export['x']) == int(g['n_export'])
    assert field_err(s.export['x'][-1], g['export_x_last']) <= 1e-9
    assert s.timing_data['total'] > 0 and set(s.timing_data) >= {'compute', 'neighbour_hood', 'time_step'}


def test_async_export_equals_blocking_export(monkeypatch):
    """The double-buffered export path fills solver.export with the same arrays as the blocking one."""
    monkeypatch.setenv("OSPH_QUIET", "1")
    from src.Solver import Solver
    from src.Methods.WCSPH import WCSPH
    from src.Kernels.CubicSpline import CubicSpline
    from src.Integrators.PEC import PEC
    runs = []
    for sync in ("1", ""):
        if sync:
            monkeypatch.setenv("OSPH_SYNC_EXPORT", sync)
        else:
            monkeypatch.delenv("OSPH_SYNC_EXPORT", raising=False)
        r0, pA = W.dam_break(10)
        method = WCSPH(height=25.0, r0=r0, rho0=1000.0, useXSPH=True, Pb=0, useSummationDensity=False)
        s = Solver(method, PEC(useXSPH=True, strict=False), CubicSpline(), 0.02, incrementalWriteout=False,
                   h=1.6 * r0, maxSettle=4, exportProperties=['x', 'y', 'p', 'vx'])
        s.addParticles(pA)
        s.setup()
        s.run()
        assert s._async_export == (not sync) and not s._export_pending
        runs.append(s)
    a, b = runs
    assert a.t_step == b.t_step
    for key in ('x', 'y', 'p', 'vx'):
        assert len(a.export[key]) == len(b.export[key]) == a.t_step
        for u, v in zip(a.export[key], b.export[key]):
            assert np.array_equal(u, v, equal_nan=True), key


def test_containment_example_runs(monkeypatch):
    """Dynamic h, XSPH off, WCSPH() without useSummationDensity: the shipped Containment call pattern."""
    monkeypatch.setenv("OSPH_QUIET", "1")
    import importlib.util, os
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("ex_containment", os.path.join(PKG, "examples", "containment.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    s = mod.main(['--nx', '20', '--duration', '0.002', '--max-settle', '5'])
    pa = s.particleArray
    assert s.t >= 0.002 and np.all(np.isfinite(pa['x'])) and np.all(np.isfinite(pa['rho']))
    fluid = pa['label'] == 0
    assert np.all(pa['h'][fluid] > 0) and pa['y'][fluid].min() > -1.0 / 20        # nothing fell through the floor
    assert len(s.export['vx']) == s.t_step


def test_kernel_objects_standalone():
    """reference test/test_kernels_cubic.py closed forms + oracle, through the device leaf entry points."""
    from src.Kernels.CubicSpline import CubicSpline
    from src.Kernels.Wendland import Wendland
    from src.Kernels.Gaussian import Gaussian
    rng = np.random.default_rng(3)
    r = rng.uniform(0, 0.9, 500); h = rng.uniform(0.2, 0.4, 500); x = rng.uniform(-1, 1, 500) * r
    r[:3] = [0.0, 1e-11, 0.5]; h[:3] = [0.25, 0.25, 0.25]
    for K, name in ((CubicSpline, 'cubic'), (Wendland, 'wendland'), (Gaussian, 'gaussian')):
        assert np.allclose(K.evaluate(r, h), O.kernel_evaluate(name, r, h), rtol=1e-13, atol=1e-300)
        assert np.allclose(K.gradient(x, r, h), O.kernel_gradient(name, x, r, h), rtol=1e-13, atol=1e-300)
    a = 10 / (7 * np.pi)
    assert np.allclose(CubicSpline.evaluate(np.array([0.5, 1.0, 1.5, 2.0, 3.0]), np.ones(5)),
                       [a * (1 - 1.5 * .25 * .75), a * .25, a * .25 * .125, 0, 0], rtol=1e-14)


def test_method_and_tools_standalone():
    from src.Methods.WCSPH import WCSPH
    from src.Tools.SolverTools import computeH, findActive, _loop
    from src.Tools.NNLinkedList import NNLinkedList
    from src.Kernels.CubicSpline import CubicSpline
    from src.Equations.TimeStep import TimeStep
    from src.Equations.KineticEnergy import KineticEnergy
    from src.Equations.TaitEOS import TaitEOS, TaitEOS_height
    g, meta, pA = load_golden('dambreak20_cubic')
    c = meta['consts']
    m = WCSPH(c['height'], c['r0'], c['rho0'], True, 0)
    assert (m.co, m.B, m.D) == (c['co'], c['B'], c['D'])
    w = O.wcsph(c['height'], c['r0'], c['rho0'], True)
    fluid = pA[pA['label'] == 0]
    assert np.allclose(m.initialize(fluid.copy())['rho'], O.initialize_density(w, fluid['y']), rtol=1e-14)
    L = O.lib()
    want = np.array([L.oracle_tait_p(7.0, c['B'], c['rho0'], float(r), int(l)) for r, l in zip(pA['rho'], pA['label'])])
    assert np.allclose(m.compute_pressure(pA), want, rtol=1e-12, atol=1e-6)
    assert np.allclose(TaitEOS(7.0, c['B'], c['rho0'], pA['rho'], pA['label']), want, rtol=1e-12, atol=1e-6)
    assert np.allclose(TaitEOS_height(c['rho0'], c['height'], c['B'], 7.0, fluid['y']),
                       O.initialize_density(w, fluid['y']), rtol=1e-14)
    assert np.allclose(computeH(1.3, len(fluid), fluid['m'], fluid['rho']), O.compute_h(1.3, fluid['m'], fluid['rho']),
                       rtol=1e-15)
    cnt, mask = findActive(0, pA)
    assert cnt == len(pA) and mask.all()
    # neighbour search object
    nn = NNLinkedList(2.0)
    nn.update(pA)
    assert nn.cell_size == g['grid'][4] and list(nn.ncells_per_dim) == [int(g['grid'][5]), int(g['grid'][6])]
    i = 7
    hh, q, r, idx = nn.near(i, pA)
    want_idx = np.sort(g['nbr_idx'][g['nbr_off'][i]:g['nbr_off'][i + 1]])
    assert idx.dtype == np.uint64 and np.array_equal(idx.astype(np.int64), want_idx)
    # _loop as a stand-alone call
    out = _loop(pA.copy(), CubicSpline.evaluate, CubicSpline.gradient, m, nn)
    for f in ('p', 'c', 'drho', 'ax', 'ay', 'xsphx', 'xsphy'):
        assert field_err(out[f], g['loop_' + f]) <= 1e-10, f
    # time step / kinetic energy on host arrays
    P = O.Particles.from_aos(fluid)
    assert TimeStep().compute(len(fluid), fluid, 0.25, 0.25) == O.timestep(P, P.fluid)
    assert KineticEnergy(len(fluid), fluid) == pytest.approx(O.kinetic_energy(P), rel=1e-13)


def test_equation_functions_standalone():
    """The reference's equation tests pointed at this package: test/test_numba_momentum.py:21-90,
    test_numba_continuity.py:12-46, test_eq_boundary.py:9-26, test_eq_courant.py:8-11 -- the functions of src.Equations run on
    the device (osph_leaf_equations / osph_leaf_courant) and are held to the golden results of the reference functions
    (tests/golden/equations_leaf.npz), the oracle leaves and the closed forms of those tests."""
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if here not in sys.path:
        sys.path.insert(0, here)
    import leaf_inputs as LI
    from src.Common import computed_dtype, particle_dtype
    from src.Equations.BoundaryForce import BoundaryForce
    from src.Equations.Continuity import Continuity
    from src.Equations.Courant import Courant
    from src.Equations.Momentum import Momentum
    from src.Equations.XSPH import XSPH
    from src.Methods.WCSPH import WCSPH
    g = np.load(os.path.join(here, "equations_leaf.npz"))

    def records(pa_fields, comp):
        pa = np.zeros(1, dtype=particle_dtype)
        for k, v in (pa_fields or {}).items():
            pa[k] = v
        c = np.zeros(len(comp), dtype=computed_dtype)
        for f in comp.dtype.names:
            c[f] = comp[f]
        return pa[0], c

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    # test_numba_momentum.py: 10 000 neighbours; the closed form of the test is its vectorised _calc_old
    pa, comp = records(*LI.momentum_reftest())
    a = Momentum(0.01, 0.0, pa, comp)
    assert rel(a, g['momentum_reftest']) <= 1e-12
    tmp = pa['p'] / pa['rho'] ** 2 + comp['p'] / comp['rho'] ** 2
    assert a[0] == pytest.approx(np.sum(-comp['m'] * tmp * comp['dw_x']), rel=1e-12)
    # test_numba_continuity.py (its `p` is an empty array)
    _, comp = records(None, LI.continuity_reftest())
    d = Continuity(np.array([]), comp)
    assert abs(d / float(g['continuity_reftest']) - 1) <= 1e-12
    assert abs(d / np.sum(comp['m'] * (comp['vx'] * comp['dw_x'] + comp['vy'] * comp['dw_y'])) - 1) <= 1e-12
    # test_eq_boundary.py: the wall on the left pushes to +x only
    pa, comp = records(None, LI.boundary_reftest())
    f = BoundaryForce(1.0, 5 * 9.81 * 1.0, 4.0, 2.0, pa, comp)
    assert f[0] > 0 and f[1] == 0.0 and rel(f, g['boundary_reftest']) <= 1e-13
    # test_eq_courant.py
    assert Courant(0.4, np.array([0.0]), np.array([1.0])) == 0.0
    assert Courant(0.4, np.array([1.0]), np.array([1.0])) == pytest.approx(0.4, abs=1e-15)
    assert Courant(0.4, np.array([2.0]), np.array([4.0])) == pytest.approx(0.2, abs=1e-15)
    # random table: mixed labels, approaching pairs, beta viscosity, general exponents, coincident wall particle
    pa_f, comp_in = LI.random_table()
    pa, comp = records(pa_f, comp_in)
    assert rel(Momentum(0.01, 0.0, pa, comp), g['random_momentum']) <= 1e-12
    assert rel(Momentum(0.3, 0.7, pa, comp), g['random_momentum_beta']) <= 1e-12
    assert rel(Momentum(0.3, 0.7, pa, comp), O.eq_momentum(0.3, 0.7, pa_f, comp_in)) <= 1e-12
    assert abs(Continuity(pa, comp) / float(g['random_continuity']) - 1) <= 1e-12
    assert rel(XSPH(0.5, pa, comp), g['random_xsph']) <= 1e-12
    assert rel(BoundaryForce(0.15, 1226.25, 4.0, 2.0, pa, comp), g['random_boundary_42']) <= 1e-12
    assert rel(BoundaryForce(0.15, 1226.25, 12.0, 6.0, pa, comp), g['random_boundary_126']) <= 1e-12
    assert Courant(0.25, comp['h'], comp['c']) == pytest.approx(float(g['courant'][3]), rel=1e-15)
    # the per-particle methods of WCSPH (reference WCSPH.py:151-203)
    m = WCSPH(1.0, 0.15, 1000.0, True, 0)
    m.alpha, m.beta = 0.3, 0.7
    acc = m.compute_acceleration(pa, comp)
    assert rel([acc[0], acc[1] + 9.81], g['random_momentum_beta']) <= 1e-12
    v = m.compute_velocity(pa, comp)
    assert v[0] == pa['vx'] and rel([v[2] - pa['vx'], v[3] - pa['vy']], g['random_xsph']) <= 1e-12
    assert abs(m.compute_density_change(pa, comp) / float(g['random_continuity']) - 1) <= 1e-12
    m.useXSPH = False; m.useSummationDensity = True
    assert m.compute_velocity(pa, comp)[2:] == [0.0, 0.0] and m.compute_density_change(pa, comp) == 0.0
    # an empty table
    empty = np.zeros(0, dtype=computed_dtype)
    assert Momentum(0.01, 0.0, pa, empty) == [0.0, 0.0] and Continuity(pa, empty) == 0.0


@pytest.mark.parametrize("name", ['tank16_gaussian', 'dambreak20_wendland'])
def test_force_evaluation_composed_from_the_plugin_objects(name):
    """One force evaluation written the way the reference composes it (SolverTools.py:143-173, and make_golden.py's
    loop_gaussian for the kernel `_loop` cannot take): nn.near -> computeProps -> method.compute_density_change /
    compute_acceleration / compute_velocity + BoundaryForce, every piece a stand-alone device leaf of this package --
    equal to the golden vectors of the reference and to the fused pair kernel."""
    from src.Equations.BoundaryForce import BoundaryForce
    from src.Kernels.Gaussian import Gaussian
    from src.Kernels.Wendland import Wendland
    from src.Methods.WCSPH import WCSPH
    from src.Tools.NNLinkedList import NNLinkedList
    from src.Tools.SolverTools import computeProps, _loop
    g, meta, pA = load_golden(name)
    c = meta['consts']
    K = {'gaussian': Gaussian, 'wendland': Wendland}[meta['kernel']]
    m = WCSPH(c['height'], c['r0'], c['rho0'], meta['useXSPH'], c['Pb'])
    pA = pA.copy()
    pA['p'] = m.compute_pressure(pA)
    pA['c'] = m.compute_speed_of_sound(pA)
    nn = NNLinkedList(2.0)
    nn.update(pA)
    fluid = np.flatnonzero(pA['label'] == 0)
    pick = fluid[::max(1, len(fluid) // 60)]
    got = {f: np.zeros(len(pick)) for f in ('drho', 'ax', 'ay', 'xsphx', 'xsphy')}
    for k, i in enumerate(pick):
        h_i, q_i, dist, near = nn.near(int(i), pA)
        comp = computeProps(int(i), pA, near, h_i, q_i, dist, K.evaluate, K.gradient)
        a = m.compute_acceleration(pA[i], comp)
        b = BoundaryForce(m.r0, m.D, m.p1, m.p2, pA[i], comp)
        v = m.compute_velocity(pA[i], comp)
        got['drho'][k] = m.compute_density_change(pA[i], comp)
        got['ax'][k] = a[0] + b[0]; got['ay'][k] = a[1] + b[1]
        got['xsphx'][k] = v[2]; got['xsphy'][k] = v[3]
    fused = _loop(pA.copy(), K.evaluate, K.gradient, m, nn)
    for f in got:
        ref = g['loop_' + f]
        scale = np.maximum(np.abs(ref[pick]), np.abs(ref).max())
        assert np.max(np.abs(got[f] - ref[pick]) / scale) <= 1e-10, (f, 'vs golden')
        assert np.max(np.abs(got[f] - fused[f][pick]) / scale) <= 1e-10, (f, 'vs fused kernel')


def test_integrator_objects_standalone():
    """reference test/test_integrators_pec.py known answers through the device kernels."""
    from src.Integrators.PEC import PEC
    from src.Integrators.Euler import Euler
    from src.Common import particle_dtype
    def some(label):
        p = np.zeros(1, dtype=particle_dtype)
        p['label'] = label; p['vx'] = 1.0; p['vy'] = 3.0; p['ax'] = 5.0; p['drho'] = 10
        return 2.0, p
    i = PEC(useXSPH=False)
    dt, p = some(0)
    p2 = i.predict(dt, p, 0)
    assert (p2['x'][0], p2['y'][0], p2['rho'][0]) == (dt * 0.5, dt * 0.5 * 3.0, dt * 0.5 * 10.0)
    p3 = i.correct(dt, p2, 0)
    assert p3['x'][0] == pytest.approx(dt * 1.0 + 0.5 * 5 * 2 * 2) and p3['y'][0] == pytest.approx(dt * 3.0)
    assert p3['rho'][0] == pytest.approx(10 * dt)
    dt, pe = some(0)
    pe = Euler().correct(dt, Euler().predict(dt, pe, 0), 0)
    for f in ('x', 'y', 'vx', 'vy', 'rho'):
        assert p3[f][0] == pytest.approx(pe[f][0])


def test_batched_pressure_probe_matches_reference_recipe():
    """osph_probe_pressure vs the IceBreak recipe evaluated with the oracle's nearPos and the leaf helpers."""
    from osph_b200 import capi
    from src.Equations.Shepard import Shepard
    from src.Equations.SummationDensity import SummationDensity
    g, meta, pA = load_golden('tank24_wendland_coupled')
    c = meta['consts']
    P = O.Particles.from_aos(pA)
    og = O.Grid(P)
    cfg = capi.make_config(c | {'useXSPH': True}, 'wendland', 'pec', capi.FP64, meta['fixed_h'], keep_h=True)
    rng = np.random.default_rng(5)
    xs = rng.uniform(0.05, 0.95, 40); ys = rng.uniform(0.05, 0.9, 40); h = 1.3 * c['r0']
    with capi.Context(cfg) as ctx:
        ctx.upload(pA); ctx.build_neighbours()
        rho, p = ctx.probe_pressure(xs, ys, h)
    for k in range(len(xs)):
        hh, q, r, idx = og.near_pos(float(xs[k]), float(ys[k]), h)
        w = O.kernel_evaluate('wendland', r, np.full_like(r, h))
        wt = Shepard(w, pA['label'][idx], pA['m'][idx], pA['rho'][idx])
        want = SummationDensity(pA['label'][idx], pA['m'][idx], wt)
        assert rho[k] == pytest.approx(want, rel=1e-12)
        assert p[k] == pytest.approx(((want / c['rho0']) ** 7 - 1) * c['B'], rel=1e-9, abs=1e-6)


def _run_coupled(rows_only):
    from src.Solver import Solver
    from src.Methods.WCSPH import WCSPH
    from src.Kernels.Wendland import Wendland
    from src.Integrators.PEC import PEC
    from src.Integrators.NewmarkBeta import NewmarkBeta
    from src.Common import ParticleType
    case = W.tank_case(16, h=1.3 / 16, useXSPH=True, seed=2, coupled_row=True, perturbed=False)
    pA = case['pA']
    nc = int((pA['label'] == ParticleType.Coupled).sum())
    calls = []

    def coupling(arr, solver):
        c = arr['label'] == ParticleType.Coupled
        rho, p = solver.probe_pressure(arr['x'][c], arr['y'][c] - 0.5 / 16, 1.3 / 16)
        assert rho.shape == (nc,) and np.all(np.isfinite(p))
        F = np.clip(p, -1e4, 1e4) * 1e-3
        a = solver.couplingIntegrator.acceleration(solver.dt, F, arr['y'][c] - y0, arr['vy'][c])
        arr['ay'][c] = a
        calls.append(float(np.abs(a).max()))
        return arr

    y0 = pA['y'][pA['label'] == ParticleType.Coupled].copy()
    nb = NewmarkBeta(0.25, 0.5, np.eye(nc) * 50.0, np.eye(nc) * 1e4, np.eye(nc) * 10.0)
    method = WCSPH(height=1.0, r0=case['r0'], rho0=1000.0, useXSPH=True, Pb=0)
    s = Solver(method, PEC(useXSPH=True, strict=False), Wendland(), 0.002, incrementalWriteout=False, h=1.3 / 16,
               maxSettle=3, coupling=coupling, couplingIntegrator=nb, couplingProperties={}, exportProperties=['y'],
               couplingRowsOnly=rows_only)
    s.addParticles(pA)
    s.setup()
    s.run()
    return s, calls, y0


def test_solver_coupling_callback_path(monkeypatch):
    """Coupled (ice-like) particles moved by a host-side NewmarkBeta integrator inside a coupling callback, the
    IceBreak call pattern (examples/IceBreak.py:109-171): predict/correct of the coupled rows on the host mirror,
    callback between compute and correct, batched device pressure probe inside the callback.  Run twice: with the
    reference's whole-array round trips, and with only the Coupled rows crossing PCIe (osph_download_rows /
    osph_upload_rows); both must give the same simulation."""
    monkeypatch.setenv("OSPH_QUIET", "1")
    from src.Common import ParticleType
    runs = []
    for rows_only in (False, True):
        s, calls, y0 = _run_coupled(rows_only)
        assert s.couplingRowsOnly == rows_only
        assert len(calls) == s.t_step and s.t >= 0.002
        out = s.particleArray
        assert np.all(np.isfinite(out['x'])) and np.all(np.isfinite(out['y']))
        c = out['label'] == ParticleType.Coupled
        assert np.any(out['y'][c] != y0)                         # the host integrator moved the coupled rows ...
        f = out['label'] == 0
        assert np.all(out['ay'][f] != 0)                          # ... and the device kept integrating the fluid
        assert s.timing_data['coupling'] > 0
        runs.append((s, calls))
    (a, ca), (b, cb) = runs
    assert a.t_step == b.t_step and np.allclose(a.dt_a, b.dt_a, rtol=1e-10, atol=0)
    assert np.allclose(ca, cb, rtol=1e-8, atol=1e-12)
    for f in ('x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay'):      # storage order differs between the modes: summation order only
        assert field_err(a.particleArray[f], b.particleArray[f]) <= 1e-9, f
    assert len(a.export['y']) == len(b.export['y']) == a.t_step
