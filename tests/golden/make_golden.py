"""Freeze golden vectors from the UNMODIFIED reference (numba path) for the WCSPH step.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The reference has no test that pins the composed step,
the neighbour order or the cell assignment, so these files are
the pin: the oracle (oracle/wcsph_oracle.c) is checked against them on CPU, and the
CUDA path is checked against both on the GPU box (where the reference is absent).

Each case stores
  aos        the input particle array as raw bytes (n x 154), state right after
             Solver.setup()-style initialisation + deterministic jitter
  meta       json: constants, kernel, integrator flags, versions
  grid       xmin xmax ymin ymax cell_size ncx ncy   (NNLinkedList after update())
  cell_ids   flat reference cell of every active particle (from heads/nexts)
  nbr_off / nbr_idx   CSR neighbour lists of the fluid rows, reference order
  loop_*     p c drho ax ay xsphx xsphy after one `_loop`
  step_*     x y vx vy rho drho ax ay after each of `nsteps` whole steps driven in
             the order of src/Solver.py:366-399, and the (dt, dt_c, dt_f) series
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))

from oracle import refshim  # noqa: E402
from osph_b200 import workloads as W  # noqa: E402

ref = refshim.load()
FLUID = 0


def ref_kernel(name):
    return {'cubic': ref.CubicSpline, 'wendland': ref.Wendland, 'gaussian': ref.Gaussian}[name]()


def cells_from_lists(nn, n):
    cid = np.full(n, -1, dtype=np.int64)
    for c in range(nn.n_cells):
        j = nn.heads[c]
        while j != -1:
            cid[j] = c
            j = nn.nexts[j]
    return cid


def neighbours(nn, pA):
    off = np.zeros(len(pA) + 1, dtype=np.int64)
    idx = []
    for i in range(len(pA)):
        off[i] = len(idx)
        if pA[i]['label'] == FLUID:
            _, _, _, nb = nn.near(i, pA)
            idx.extend(int(v) for v in nb)
    off[len(pA)] = len(idx)
    return off, np.asarray(idx, dtype=np.int32)


def loop_gaussian(pA, method, nn):
    """`_loop` cannot take the Gaussian ufuncs (numba typing); compose it from the reference's pieces."""
    G = ref.Gaussian
    pA['p'] = method.compute_pressure(pA)
    pA['c'] = method.compute_speed_of_sound(pA)
    for i in range(len(pA)):
        if pA[i]['label'] != FLUID:
            continue
        h_i, q_i, dist, near = nn.near(i, pA)
        if len(near) == 0:
            continue
        comp = ref._assignProps(i, pA, near, h_i, q_i, dist)
        comp['w'] = G.evaluate(comp['r'], comp['h'])
        comp['dw_x'] = G.gradient(comp['x'], comp['r'], comp['h'])
        comp['dw_y'] = G.gradient(comp['y'], comp['r'], comp['h'])
        pA[i]['drho'] = method.compute_density_change(pA[i], comp)
        a = method.compute_acceleration(pA[i], comp)
        b = ref.BoundaryForce(method.r0, method.D, method.p1, method.p2, pA[i], comp)
        pA[i]['ax'] = a[0] + b[0]
        pA[i]['ay'] = a[1] + b[1]
        v = method.compute_velocity(pA[i], comp)
        pA[i]['vx'], pA[i]['vy'], pA[i]['xsphx'], pA[i]['xsphy'] = v
    return pA


def run_loop(pA, kernel, method, nn):
    if kernel == 'gaussian':
        return loop_gaussian(pA, method, nn)
    k = ref_kernel(kernel)
    return ref._loop(pA, k.evaluate, k.gradient, method, nn)


def make_case(name, case, kernel, useXSPH, damping, nsteps, strict=False, scale=2.0, summation=False):
    pA0 = case['pA'].copy()
    c = case['consts']
    n = len(pA0)
    fixed_h = case['h']
    method = ref.WCSPH(c['height'], c['r0'], c['rho0'], useXSPH, c['Pb'], summation)
    assert method.co == c['co'] and method.B == c['B'] and method.D == c['D']
    integ = ref.PEC(useXSPH, strict)
    out = dict(aos=np.frombuffer(pA0.tobytes(), dtype=np.uint8).reshape(n, 154).copy())

    # --- neighbour structure + one force evaluation on the input state ------------------------
    pA = pA0.copy().view(ref.particle_dtype)
    nn = ref.NNLinkedList(scale)
    nn.update(pA)
    out['grid'] = np.array([nn.xmin, nn.xmax, nn.ymin, nn.ymax, nn.cell_size,
                            nn.ncells_per_dim[0], nn.ncells_per_dim[1]], dtype=np.float64)
    out['cell_ids'] = cells_from_lists(nn, n)
    out['nbr_off'], out['nbr_idx'] = neighbours(nn, pA)
    pA = run_loop(pA, kernel, method, nn)
    for f in ('p', 'c', 'drho', 'ax', 'ay', 'xsphx', 'xsphy'):
        out['loop_' + f] = pA[f].copy()

    # --- whole steps, same order as Solver.run() ---------------------------------------------
    pA = pA0.copy().view(ref.particle_dtype)
    fi = pA['label'] == FLUID
    nf = int(fi.sum())
    dts = []
    cols = ('x', 'y', 'vx', 'vy', 'rho', 'drho', 'ax', 'ay', 'xsphx', 'xsphy', 'p', 'h')
    hist = {f: [] for f in cols}
    for _ in range(nsteps):
        m, dc, df = ref.TimeStep().compute(nf, pA[fi], 0.25, 0.25)
        dts.append((m, dc, df))
        pA[fi] = integ.predict(m, pA[fi], damping)
        nn.update(pA)
        if fixed_h is None:
            pA['h'][fi] = ref.computeH(1.3, nf, pA[fi]['m'], pA[fi]['rho'])
        else:
            pA['h'][fi] = fixed_h
        pA = run_loop(pA, kernel, method, nn)
        pA[fi] = integ.correct(m, pA[fi], damping)
        for f in cols:
            hist[f].append(pA[f].copy())
    out['dts'] = np.asarray(dts, dtype=np.float64).reshape(nsteps, 3)
    for f in cols:
        out['step_' + f] = np.stack(hist[f]) if nsteps else np.zeros((0, n))
    import numba
    out['meta'] = np.frombuffer(json.dumps(dict(
        name=name, kernel=kernel, useXSPH=bool(useXSPH), strict=bool(strict), damping=damping,
        nsteps=nsteps, fixed_h=fixed_h, scale=scale, consts=c, n=n, n_fluid=nf, summation=bool(summation),
        numba=numba.__version__, numpy=np.__version__)).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s n=%5d fluid=%5d pairs=%7d grid=%dx%d cs=%g  -> %.0f kB' % (
        name, n, nf, len(out['nbr_idx']), out['grid'][5], out['grid'][6], out['grid'][4],
        os.path.getsize(path) / 1e3))


def solver_run_case(name, N, kernel, duration, maxSettle):
    """The reference's own Solver.setup()/run() on the dam break, end state only."""
    r0, pA = W.dam_break(N)
    pA = pA.view(ref.particle_dtype)
    k = ref_kernel(kernel)
    method = ref.WCSPH(height=25.0, r0=r0, rho0=1000.0, useXSPH=True, Pb=0, useSummationDensity=False)
    integ = ref.PEC(useXSPH=True, strict=False)
    S = ref.Solver
    S.particleArray = None; S.data = []; S.export = {}; S.dt_a = []; S.dt_c = []; S.dt_f = []
    s = S(method, integ, k, duration, incrementalWriteout=False, h=1.6 * r0, maxSettle=maxSettle)
    s.addParticles(pA.copy())
    s.setup()
    s.run()
    out = dict(
        final=np.frombuffer(s.particleArray.tobytes(), dtype=np.uint8).reshape(len(pA), 154).copy(),
        dt_a=np.asarray(s.dt_a), dt_c=np.asarray(s.dt_c), dt_f=np.asarray(s.dt_f),
        t_step=np.int64(s.t_step), t=np.float64(s.t), settleTime=np.float64(s.settleTime),
        export_x_last=np.asarray(s.export['x'][-1]), n_export=np.int64(len(s.export['x'])),
        meta=np.frombuffer(json.dumps(dict(name=name, N=N, kernel=kernel, duration=duration,
                                           maxSettle=maxSettle, hfac=1.6)).encode(), dtype=np.uint8))
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s steps=%d t=%.5f settle=%.5f -> %.0f kB' % (name, s.t_step, s.t, s.settleTime,
                                                            os.path.getsize(path) / 1e3))


def verlet_case(name='verlet_leaf'):
    """The reference's Verlet integrator (src/Integrators/Verlet.py:28-55) on a random fluid block: state after
    predict and after correct, with and without XSPH.  Pins the oracle's Verlet (the GPU path is checked against it)."""
    rng = np.random.default_rng(11)
    n = 96
    pA = np.zeros(n, dtype=ref.particle_dtype)
    for f in ('x', 'y', 'vx', 'vy', 'ax', 'ay', 'xsphx', 'xsphy', 'drho'):
        pA[f] = rng.normal(size=n) * {'ax': 9.0, 'ay': 9.0, 'drho': 40.0}.get(f, 1.5)
    pA['rho'] = 1000.0 + rng.normal(size=n)
    pA['m'] = 1.0
    dt, damping = 3.7e-4, 0.05
    out = dict(aos=np.frombuffer(pA.tobytes(), dtype=np.uint8).reshape(n, 154).copy(), dt=np.float64(dt))
    for useXSPH in (True, False):
        integ = ref.Verlet(useXSPH)
        assert integ.isMultiStage() is False
        q = integ.predict(dt, pA.copy(), damping)
        tag = 'xsph' if useXSPH else 'raw'
        out['pred_%s' % tag] = np.frombuffer(q.tobytes(), dtype=np.uint8).reshape(n, 154).copy()
        q = integ.correct(dt, q, damping)
        out['corr_%s' % tag] = np.frombuffer(q.tobytes(), dtype=np.uint8).reshape(n, 154).copy()
    import numba
    out['meta'] = np.frombuffer(json.dumps(dict(name=name, n=n, dt=dt, damping=damping, numba=numba.__version__,
                                                numpy=np.__version__)).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s n=%d -> %.0f kB' % (name, n, os.path.getsize(path) / 1e3))


def equations_leaf_case(name='equations_leaf'):
    """The reference's equations called one by one (Momentum, Continuity, XSPH, BoundaryForce, Courant) on the synthetic
    neighbour tables of its own equation tests and on a random table (tests/golden/leaf_inputs.py).  Only the results are
    stored; the inputs are regenerated from the same module by the test."""
    import leaf_inputs as LI
    import importlib.util
    spec = importlib.util.spec_from_file_location("_osph_reference_courant",
                                                  os.path.join(refshim.REFERENCE_ROOT, "src", "Equations", "Courant.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                       # the module only needs numba
    Courant = mod.Courant

    def to_ref(pa_fields, comp):
        pa = np.zeros(1, dtype=ref.particle_dtype)
        for k, v in (pa_fields or {}).items():
            pa[k] = v
        c = np.zeros(len(comp), dtype=ref.computed_dtype)
        for f in comp.dtype.names:
            c[f] = comp[f]
        return pa[0], c

    out = {}
    pa_f, comp = LI.momentum_reftest()
    pa, c = to_ref(pa_f, comp)
    out['momentum_reftest'] = np.asarray(ref.Momentum(0.01, 0.0, pa, c), dtype=np.float64)
    pa, c = to_ref(None, LI.continuity_reftest())
    out['continuity_reftest'] = np.float64(ref.Continuity(pa, c))
    pa, c = to_ref(None, LI.boundary_reftest())
    out['boundary_reftest'] = np.asarray(ref.BoundaryForce(1.0, 5 * 9.81 * 1.0, 4.0, 2.0, pa, c), dtype=np.float64)
    pa_f, comp = LI.random_table()
    pa, c = to_ref(pa_f, comp)
    out['random_momentum'] = np.asarray(ref.Momentum(0.01, 0.0, pa, c), dtype=np.float64)
    out['random_momentum_beta'] = np.asarray(ref.Momentum(0.3, 0.7, pa, c), dtype=np.float64)
    out['random_continuity'] = np.float64(ref.Continuity(pa, c))
    out['random_xsph'] = np.asarray(ref.XSPH(0.5, pa, c), dtype=np.float64)
    out['random_boundary_42'] = np.asarray(ref.BoundaryForce(0.15, 1226.25, 4.0, 2.0, pa, c), dtype=np.float64)
    out['random_boundary_126'] = np.asarray(ref.BoundaryForce(0.15, 1226.25, 12.0, 6.0, pa, c), dtype=np.float64)
    out['courant'] = np.asarray([Courant(0.4, np.array([0.0]), np.array([1.0])), Courant(0.4, np.array([1.0]), np.array([1.0])),
                                 Courant(0.4, np.array([2.0]), np.array([4.0])),
                                 Courant(0.25, comp['h'], comp['c'])], dtype=np.float64)
    import numba
    out['meta'] = np.frombuffer(json.dumps(dict(name=name, numba=numba.__version__, numpy=np.__version__)).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s -> %.1f kB' % (name, os.path.getsize(path) / 1e3))
    for k, v in out.items():
        if k != 'meta':
            print('   ', k, v)


def main():
    # regime A: 3h > reference cell (1.0): the 3x3 coarse walk truncates the neighbourhood
    make_case('dambreak20_wendland', W.dam_break_case(20), 'wendland', True, 0.05, 3)
    make_case('dambreak20_cubic', W.dam_break_case(20, seed=1), 'cubic', True, 0.0, 3)
    # regime B: 3h < cell; dynamic h (Containment-like: h=None, XSPH off, cubic)
    make_case('tank30_cubic_dynh', W.tank_case(30, h=None, useXSPH=False), 'cubic', False, 0.05, 3)
    # coupled (ice-like) row with mass, XSPH on, Wendland, strict density clamp
    make_case('tank24_wendland_coupled', W.tank_case(24, h=1.6 / 24, useXSPH=True, coupled_row=True, seed=2),
              'wendland', True, 0.0, 2, strict=True)
    # Gaussian: composed oracle, one force evaluation + one step
    make_case('tank16_gaussian', W.tank_case(16, h=1.3 / 16, useXSPH=True, seed=3), 'gaussian', True, 0.05, 1)
    # no boundary particles: every h > 0, so the reference cell is 2*hmin instead of the 1.0 fallback
    blk = W.tank_case(20, h=1.5 / 20, useXSPH=True, seed=4)
    keep = blk['pA']['label'] == FLUID
    blk['pA'] = blk['pA'][keep]
    make_case('block20_cubic_nobnd', blk, 'cubic', True, 0.0, 2)
    # optional summation-density branch of _loop (off in every shipped example)
    make_case('tank16_cubic_sumdens', W.tank_case(16, h=1.3 / 16, useXSPH=True, seed=5), 'cubic', True, 0.0, 2,
              summation=True)
    # the reference Solver end to end (settle -> gate removal -> time stepping)
    solver_run_case('solver_dambreak12_wendland', 12, 'wendland', 0.03, 6)
    verlet_case()
    equations_leaf_case()


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'leaf':         # regenerate only the leaf-equation vectors
        equations_leaf_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'sumdens':      # regenerate only the newest case
        make_case('tank16_cubic_sumdens', W.tank_case(16, h=1.3 / 16, useXSPH=True, seed=5), 'cubic', True, 0.0, 2,
                  summation=True)
    elif len(sys.argv) > 1 and sys.argv[1] == 'verlet':
        verlet_case()
    else:
        main()
