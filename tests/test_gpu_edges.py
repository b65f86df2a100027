"""GPU: rarely taken paths of the device code against the CPU oracle -- general Lennard-Jones exponents and
beta-viscosity (the pow() branch of the pair kernel), coincident particles (the r = 0 guards of the reference),
clusters far denser than the candidate list, a domain that outgrows its cell table, parked non-finite particles,
empty and fluid-free inputs.  The same functions run on the CPU under the SIMT emulator (tests/test_emu_kernels.py).
"""
import numpy as np
import pytest

from conftest import field_err
from oracle import oracle as O
from osph_b200 import capi
from osph_b200 import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-10
STATE_FIELDS = ('x', 'y', 'vx', 'vy', 'rho', 'drho', 'ax', 'ay', 'xsphx', 'xsphy', 'p', 'h')


def _oracle_params(c, **over):
    w = O.wcsph(c['height'], c['r0'], c['rho0'], True)
    for k, v in over.items():
        setattr(w, k, v)
    return w


def _compare_steps(case, kernel, steps, consts, w, damping=0.05, fixed_dt=None):
    pA = case['pA']
    P = O.Particles.from_aos(pA)
    cfg = capi.make_config(consts, kernel, 'pec', capi.FP64, case['h'])
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        for s in range(steps):
            dt3, _ = O.step(P, w, kernel, 'pec', True, False, damping, case['h'], fixed_dt)
            ctx.step(1, fixed_dt, damping)
            cols = ctx.download_fields(list(STATE_FIELDS))
            for f in STATE_FIELDS:
                assert field_err(cols[f], getattr(P, f)) <= TOL, (s, f, field_err(cols[f], getattr(P, f)))
            assert np.allclose(ctx.dt_log()[-1], dt3, rtol=1e-10)
        return ctx.sync()


@pytest.mark.parametrize("kernel", ['cubic', 'gaussian'])
def test_general_lennard_jones_exponents_and_beta_viscosity(kernel):
    """p1, p2 != (4, 2) takes the pow() branch of the wall force; beta != 0 the quadratic viscosity term."""
    case = W.dam_break_case(30, seed=3)
    consts = dict(case['consts'], p1=6.0, p2=3.0, beta=0.7, alpha=0.2, epsilon=0.3)
    w = _oracle_params(case['consts'], p1=6.0, p2=3.0, beta=0.7, alpha=0.2, epsilon=0.3)
    assert _compare_steps(case, kernel, 2, consts, w) == 0


@pytest.mark.parametrize("kernel", ['wendland', 'cubic'])
def test_background_pressure_and_other_gas_constants(kernel):
    """Pb != 0 (added to the pressure of EVERY label, WCSPH.py:143-149, so the wall rows carry Pb and the fluid rows
    Tait + Pb), gamma, alpha and epsilon away from their defaults, no damping."""
    case = W.tank_case(24, h=1.5 / 24, useXSPH=True, seed=9)
    over = dict(Pb=3.5e3, gamma=5.0, alpha=0.05, epsilon=0.25)
    consts = dict(case['consts'], **over)
    consts['B'] = consts['co'] ** 2 * consts['rho0'] / consts['gamma']           # TaitEOS_B with the changed gamma
    w = _oracle_params(case['consts'], **dict(over, B=consts['B']))
    assert _compare_steps(case, kernel, 2, consts, w, damping=0.0) == 0


@pytest.mark.parametrize("scale,fluid_only", [(3.0, True), (1.0, True), (3.0, False)])
def test_neighbour_scale_other_than_two(scale, fluid_only):
    """NNLinkedList(scale) with scale != 2 (the reference's test_linked_list.py uses 3): reference cell = scale * min h.
    scale = 1 makes the cell SMALLER than the 3 h_ij radius, so the 3x3 cell walk truncates the q <= 3 set and the cell
    adjacency decides membership; with boundary rows (h = 0) the cell is the 1.0 fallback whatever the scale."""
    case = W.tank_case(20, h=1.5 / 20, useXSPH=True, seed=6)
    pA = case['pA'] if not fluid_only else case['pA'][case['pA']['label'] == 0]
    P = O.Particles.from_aos(pA)
    grid = O.Grid(P, scale)
    cfg = capi.make_config(case['consts'], 'cubic', 'pec', capi.FP64, None, keep_h=True)
    cfg.nn_scale = scale
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        ctx.build_neighbours()
        g, cells = ctx.cells()
        want = grid.params
        assert [g['xmin'], g['xmax'], g['ymin'], g['ymax'], g['cell_size'], g['ncx'], g['ncy']] == \
               [want[k] for k in ('xmin', 'xmax', 'ymin', 'ymax', 'cell_size', 'ncx', 'ncy')]
        assert np.array_equal(cells, grid.cell_ids())
        off, idx = ctx.neighbours_csr()
        woff, widx = grid.neighbours_csr()
        assert np.array_equal(off, woff)
        for i in range(len(off) - 1):
            assert np.array_equal(np.sort(idx[off[i]:off[i + 1]]), np.sort(widx[woff[i]:woff[i + 1]])), i
        # one force evaluation on that neighbour structure
        w = _oracle_params(case['consts'])
        Q = P.copy()
        O.loop(Q, w, grid, 'cubic')
        ctx.compute()
        cols = ctx.download_fields(['drho', 'ax', 'ay', 'xsphx', 'xsphy'])
        for f in cols:
            assert field_err(cols[f], getattr(Q, f)) <= TOL, (f, field_err(cols[f], getattr(Q, f)))
        assert ctx.sync() == 0


def test_domain_far_from_the_origin():
    """The same dam break translated by (12 345.678, -9 876.5): the grid origin is the per-step minimum, so cell ids and
    neighbour sets still follow the reference bit for bit, the FP64 fields agree with the oracle, and the FP32 mode (which
    subtracts a per-CTA anchor in double before rounding to float) stays as close to FP64 as it is at the origin."""
    case = W.dam_break_case(30, seed=8)
    case['pA'] = case['pA'].copy()
    case['pA']['x'] += 12345.678; case['pA']['y'] -= 9876.5
    pA = case['pA']
    c = dict(case['consts'])
    # the hydrostatic reference height travels with the block (H - y enters only through the initial density, already set)
    w = _oracle_params(c)
    P = O.Particles.from_aos(pA)
    grid = O.Grid(P, 2.0)
    cfg = capi.make_config(c, 'wendland', 'pec', capi.FP64, case['h'])
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        ctx.build_neighbours()
        g, cells = ctx.cells()
        assert np.array_equal(cells, grid.cell_ids())
        off, idx = ctx.neighbours_csr()
        woff, widx = grid.neighbours_csr()
        assert np.array_equal(off, woff)
        assert all(np.array_equal(np.sort(idx[off[i]:off[i + 1]]), np.sort(widx[woff[i]:woff[i + 1]])) for i in range(len(off) - 1))
    assert _compare_steps(case, 'wendland', 2, c, w) == 0
    out = {}
    for prec in (capi.FP64, capi.FP32):
        with capi.Context(capi.make_config(c, 'wendland', 'pec', prec, case['h'])) as ctx:
            ctx.upload(pA)
            ctx.step(3, None, 0.05)
            out[prec] = ctx.download_fields(['x', 'y', 'rho', 'ax', 'ay'])
    r0 = case['r0']
    assert np.max(np.abs(out[capi.FP32]['x'] - out[capi.FP64]['x'])) < 1e-4 * r0
    assert np.max(np.abs(out[capi.FP32]['y'] - out[capi.FP64]['y'])) < 1e-4 * r0
    assert np.max(np.abs(out[capi.FP32]['rho'] / out[capi.FP64]['rho'].clip(1.0) - out[capi.FP64]['rho'] / out[capi.FP64]['rho'].clip(1.0))) < 1e-5
    assert field_err(out[capi.FP32]['ax'], out[capi.FP64]['ax']) < 1e-3 and field_err(out[capi.FP32]['ay'], out[capi.FP64]['ay']) < 1e-3


def test_coincident_particles_follow_the_reference_guards():
    """Two fluid particles on one point (gradient zeroed below r = 1e-10, CubicSpline.py:47-49) and a fluid particle on
    top of a wall particle (Lennard-Jones only for r > 1e-12, BoundaryForce.py:26) must not produce NaN and must
    agree with the oracle."""
    case = W.dam_break_case(24, seed=4)
    pA = case['pA']
    fl = np.flatnonzero(pA['label'] == 0)
    bd = np.flatnonzero(pA['label'] == 1)
    a, b, c = fl[10], fl[11], fl[40]
    pA['x'][b], pA['y'][b] = pA['x'][a], pA['y'][a]
    pA['x'][c], pA['y'][c] = pA['x'][bd[5]], pA['y'][bd[5]]
    w = _oracle_params(case['consts'])
    # one evaluation with a fixed dt: the second step of such a state is dominated by the 1/r^4 wall force
    status = _compare_steps(case, 'cubic', 1, case['consts'], w, fixed_dt=1e-5)
    assert status == 0


def test_cluster_denser_than_the_candidate_list():
    """A few hundred particles inside one kernel radius: per-thread lists overflow many times (flush inside the scan),
    one acceleration cell holds more particles than a CTA."""
    case = W.dam_break_case(20, seed=8)
    pA = case['pA']
    fl = np.flatnonzero(pA['label'] == 0)
    rng = np.random.default_rng(1)
    k = 330
    centre = (pA['x'][fl[200]], pA['y'][fl[200]])
    pick = fl[:k]
    pA['x'][pick] = centre[0] + 0.4 * case['h'] * rng.uniform(-1, 1, k)
    pA['y'][pick] = centre[1] + 0.4 * case['h'] * rng.uniform(-1, 1, k)
    P = O.Particles.from_aos(pA)
    w = _oracle_params(case['consts'])
    cfg = capi.make_config(case['consts'], 'wendland', 'pec', capi.FP64, case['h'])
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        ctx.compute()
        og = O.Grid(P)
        O.loop(P, w, og, 'wendland')
        off, idx = ctx.neighbours_csr()
        ooff, oidx = og.neighbours_csr()
        assert np.array_equal(off, ooff) and int(np.max(np.diff(off))) >= k
        assert all(np.array_equal(np.sort(idx[off[i]:off[i + 1]]), np.sort(oidx[off[i]:off[i + 1]])) for i in range(len(off) - 1))
        cols = ctx.download_fields(['drho', 'ax', 'ay', 'xsphx', 'xsphy', 'p'])
        for f in cols:
            assert field_err(cols[f], getattr(P, f)) <= TOL, f


def test_domain_outgrowing_the_cell_table_coarsens_then_resizes():
    """The cell table is sized at the first build.  When a particle later flies far away the device coarsens its
    acceleration grid in place (status bit GRID_COARSE), results stay those of the oracle, and the next build after the
    caller read the status re-sizes the table."""
    case = W.dam_break_case(100, seed=9)          # fine-cell regime: pair radius 0.8 m < reference cell 1 m
    pA, c = case['pA'], case['consts']
    w = _oracle_params(c)
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    P = O.Particles.from_aos(pA)
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        O.step(P, w, 'cubic', 'pec', True, False, 0.05, case['h'])
        ctx.step(1, None, 0.05)
        assert ctx.sync() == 0
        # one fluid particle leaves: the bounding box grows ~40-fold in x and y
        i = int(np.flatnonzero(pA['label'] == 0)[-1])
        x = ctx.download_fields(['x', 'y'])
        x['x'][i] = 7000.0; x['y'][i] = 1200.0
        ctx.upload_fields(x)
        P.x[i] = 7000.0; P.y[i] = 1200.0
        seen = 0
        for s in range(3):
            O.step(P, w, 'cubic', 'pec', True, False, 0.05, case['h'])
            ctx.step(1, None, 0.05)
            cols = ctx.download_fields(list(STATE_FIELDS))
            for f in STATE_FIELDS:
                assert field_err(cols[f], getattr(P, f)) <= TOL, (s, f)
            st = ctx.sync()
            assert st & ~capi.S_GRID_COARSE == 0
            seen |= st
            if s > 0:
                assert st == 0, "the table was re-sized after the status was read: no coarsening any more"
        assert seen & capi.S_GRID_COARSE


def test_reference_grid_outgrowing_the_table_falls_back_exactly_then_resizes():
    """Where the acceleration grid IS the reference grid (pair radius >= reference cell, small N) it cannot be
    coarsened.  A domain that outgrows the table runs on a one-cell fallback grid (every particle a candidate,
    membership by the stored reference cell ids and the distance): results stay those of the oracle, the status
    reports it, and the next build re-sizes the table."""
    case = W.dam_break_case(40, seed=9)
    pA, c = case['pA'], case['consts']
    w = _oracle_params(c)
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    P = O.Particles.from_aos(pA)
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        O.step(P, w, 'cubic', 'pec', True, False, 0.05, case['h'])
        ctx.step(1, None, 0.05)
        assert ctx.sync() == 0
        i = int(np.flatnonzero(pA['label'] == 0)[-1])
        x = ctx.download_fields(['x', 'y'])
        x['x'][i] = 7000.0; x['y'][i] = 1200.0
        ctx.upload_fields(x)
        P.x[i] = 7000.0; P.y[i] = 1200.0
        for s in range(3):
            O.step(P, w, 'cubic', 'pec', True, False, 0.05, case['h'])
            ctx.step(1, None, 0.05)
            cols = ctx.download_fields(list(STATE_FIELDS))
            for f in STATE_FIELDS:
                assert field_err(cols[f], getattr(P, f)) <= TOL, (s, f)
            if s == 0:
                off, idx = ctx.neighbours_csr()                      # the validation query works on the fallback grid too
                ooff, oidx = O.Grid(P).neighbours_csr()
                assert np.array_equal(off, ooff)
            st = ctx.sync()
            assert st == (capi.S_GRID_COARSE if s == 0 else 0), (s, st)


def test_non_finite_particle_is_parked_and_reported():
    """A particle whose position is NaN is reported (status NONFINITE), finds nothing and is found by nobody: every
    other particle gets the result of the run without it."""
    case = W.dam_break_case(24, seed=10)
    pA, c = case['pA'], case['consts']
    fl = np.flatnonzero(pA['label'] == 0)
    bad = int(fl[123])
    with_nan = pA.copy()
    with_nan['x'][bad] = np.nan
    without = pA.copy()
    without['deleted'][bad] = True
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    outs = []
    for arr in (with_nan, without):
        with capi.Context(cfg) as ctx:
            ctx.upload(arr)
            ctx.compute()
            outs.append((ctx.download(arr.copy()), ctx.sync()))
    (a, sa), (b, sb) = outs
    assert sa & capi.S_NONFINITE and sb == 0
    keep = np.ones(len(pA), bool); keep[bad] = False
    for f in ('drho', 'ax', 'ay', 'xsphx', 'xsphy', 'p'):
        assert np.array_equal(a[f][keep], b[f][keep]), f


def test_empty_and_fluid_free_inputs():
    case = W.dam_break_case(12, seed=2)
    pA, c = case['pA'], case['consts']
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    with capi.Context(cfg) as ctx:
        none = pA.copy(); none['deleted'] = True
        ctx.upload(none)
        assert ctx.num_active == 0
        with pytest.raises(capi.OsphError):
            ctx.step(1, None, 0.0)
        back = ctx.download(none.copy())
        assert back.tobytes() == none.tobytes()                       # deleted rows come back verbatim
        walls = pA[pA['label'] != 0].copy()
        ctx.upload(walls)
        assert ctx.num_active == len(walls) and ctx.num_fluid == 0
        with pytest.raises(capi.OsphError):
            ctx.step(1, None, 0.0)                                    # no fluid: no time step (TimeStep.py:58-91)
        ctx.step(2, 1e-4, 0.0)                                        # a fixed dt still steps: nothing moves
        out = ctx.download(walls.copy())
        for f in ('x', 'y', 'vx', 'vy', 'rho'):
            assert np.array_equal(out[f], walls[f]), f
        off, idx = ctx.neighbours_csr()
        assert int(off[-1]) == 0                                      # wall particles are never queried


def test_fp32_mode_uses_the_reference_cell_rule_too():
    """Membership of a pair is decided by the reference cells (3 x 3) AND the distance.  Candidates normally come from
    adjacent cells by construction; on the one-cell fallback grid (reference grid outgrew the table) and for irregularly
    binned particles they do not, and the cell test has to bind.  The float instantiation applies it like the double one
    (it used to decide by distance alone there): the two modes agree to float rounding, far below one pair's share."""
    case = W.dam_break_case(40, seed=9)            # pair radius 2 m > reference cell 1 m: support reaches two cells away
    pA, c = case['pA'], case['consts']
    outs = {}
    for prec in (capi.FP64, capi.FP32):
        cfg = capi.make_config(c, 'cubic', 'pec', prec, case['h'])
        with capi.Context(cfg) as ctx:
            ctx.upload(pA)
            ctx.step(1, None, 0.05)
            assert ctx.sync() == 0
            i = int(np.flatnonzero(pA['label'] == 0)[-1])
            x = ctx.download_fields(['x', 'y'])
            x['x'][i] = 7000.0; x['y'][i] = 1200.0                 # the reference grid no longer fits the cell table
            ctx.upload_fields(x)
            ctx.compute()
            outs[prec] = ctx.download_fields(['drho', 'ax', 'ay', 'xsphx', 'xsphy'])
            assert ctx.sync() == capi.S_GRID_COARSE                # ran on the one-cell fallback
    for f in outs[capi.FP64]:
        assert field_err(outs[capi.FP32][f], outs[capi.FP64][f]) <= 5e-4, f
