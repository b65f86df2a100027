"""GPU: the CUDA path (through the C ABI) against the golden vectors frozen from the reference and
against the CPU oracle on the same seeded inputs.

Bars (north star): cell assignment and neighbour SETS bit-exact; FP64 fields within 1e-10
(norm-wise, conftest.field_err) per step; dt within 1e-10 relative.
"""
import numpy as np
import pytest

from conftest import STEP_CASES, field_err, load_golden
from oracle import oracle as O
from osph_b200 import capi
from osph_b200 import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-10
LOOP_FIELDS = ('p', 'c', 'drho', 'ax', 'ay', 'xsphx', 'xsphy')
STATE_FIELDS = ('x', 'y', 'vx', 'vy', 'rho', 'drho', 'ax', 'ay', 'xsphx', 'xsphy', 'p', 'h')


def _ctx(meta, precision=capi.FP64, keep_h=False, integrator='pec', kernel=None, reorder_every=0):
    cfg = capi.make_config(meta['consts'] | {'useXSPH': meta['useXSPH'], 'useSummationDensity': meta.get('summation', False)},
                           kernel or meta['kernel'], integrator,
                           precision, meta['fixed_h'], strict=meta['strict'], keep_h=keep_h,
                           reorder_every=reorder_every)
    return capi.Context(cfg)


def _sorted_lists(off, idx):
    return [np.sort(idx[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", STEP_CASES)
def test_cells_and_neighbour_sets_bit_exact(name):
    g, meta, pA = load_golden(name)
    with _ctx(meta, keep_h=True) as ctx:
        ctx.upload(pA)
        ctx.build_neighbours()
        grid, cells = ctx.cells()
        assert [grid['xmin'], grid['xmax'], grid['ymin'], grid['ymax'], grid['cell_size']] == list(g['grid'][:5])
        assert (grid['ncx'], grid['ncy']) == (int(g['grid'][5]), int(g['grid'][6]))
        assert np.array_equal(cells, g['cell_ids'])
        off, idx = ctx.neighbours_csr()
        assert np.array_equal(off, g['nbr_off'])
        want = _sorted_lists(g['nbr_off'], g['nbr_idx'].astype(np.int64))
        got = _sorted_lists(off, idx)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        assert ctx.sync() == 0


@pytest.mark.parametrize("name", STEP_CASES)
def test_loop_fields_vs_golden(name):
    g, meta, pA = load_golden(name)
    with _ctx(meta, keep_h=True) as ctx:
        ctx.upload(pA)
        ctx.compute()
        out = ctx.download(pA.copy())
        for f in LOOP_FIELDS:
            assert field_err(out[f], g['loop_' + f]) <= TOL, f
        # columns the loop does not write come back bit-identical
        for f in ('x', 'y', 'vx', 'vy', 'm', 'h', 'x0', 'rho0') + (() if meta.get('summation') else ('rho',)):
            assert np.array_equal(out[f], pA[f]), f
        if meta.get('summation'):        # the density written by the summation pass: against the oracle's _loop
            P = O.Particles.from_aos(pA)
            c = meta['consts']
            O.loop(P, O.wcsph(c['height'], c['r0'], c['rho0'], meta['useXSPH'], c['Pb'], True), O.Grid(P, meta['scale']),
                   meta['kernel'])
            assert field_err(out['rho'], P.rho) <= TOL
        assert np.array_equal(out['label'], pA['label']) and np.array_equal(out['deleted'], pA['deleted'])


@pytest.mark.parametrize("name", STEP_CASES)
def test_whole_steps_vs_golden(name):
    g, meta, pA = load_golden(name)
    with _ctx(meta) as ctx:
        ctx.upload(pA)
        for s in range(meta['nsteps']):
            dt, dc, df = ctx.timestep()
            assert np.allclose([dt, dc, df], g['dts'][s], rtol=1e-10, atol=0)
            ctx.predict(dt, meta['damping'])
            ctx.build_neighbours()
            ctx.compute()
            ctx.correct(dt, meta['damping'])
            cols = ctx.download_fields(list(STATE_FIELDS))
            for f in STATE_FIELDS:
                assert field_err(cols[f], g['step_' + f][s]) <= TOL, (s, f)
        assert ctx.sync() == 0


@pytest.mark.parametrize("name", ['dambreak20_wendland', 'tank30_cubic_dynh'])
def test_fused_loop_equals_explicit_calls(name):
    g, meta, pA = load_golden(name)
    n = 4
    with _ctx(meta) as a, _ctx(meta) as b:
        a.upload(pA); b.upload(pA)
        dts = []
        for _ in range(n):
            dt = a.timestep(); dts.append(dt)
            a.predict(dt[0], meta['damping']); a.build_neighbours(); a.compute(); a.correct(dt[0], meta['damping'])
        b.step(n, None, meta['damping'])
        A = a.download(pA.copy()); B = b.download(pA.copy())
        assert A.tobytes() == B.tobytes()
        assert np.array_equal(b.dt_log(), np.asarray(dts))


@pytest.mark.parametrize("fixed_dt", [None, 2e-4])
def test_multi_step_call_equals_single_step_calls(fixed_dt):
    """osph_step(n) fuses the corrector of step k with the predictor of step k+1 and takes the dt reductions from the
    predictor / pair kernel; the result must be the bytes of n separate osph_step(1) calls, however the calls are cut."""
    g, meta, pA = load_golden('tank30_cubic_dynh')
    with _ctx(meta) as a, _ctx(meta) as b, _ctx(meta) as c:
        for ctx in (a, b, c):
            ctx.upload(pA)
        for _ in range(6):
            a.step(1, fixed_dt, 0.05)
        b.step(6, fixed_dt, 0.05)
        c.step(2, fixed_dt, 0.05); c.step(1, fixed_dt, 0.05); c.step(3, fixed_dt, 0.05)
        A, B, C = (ctx.download(pA.copy()) for ctx in (a, b, c))
        assert A.tobytes() == B.tobytes() == C.tobytes()
        la, lb, lc = a.dt_log(), b.dt_log(), c.dt_log()
        assert np.array_equal(la, lb) and np.array_equal(la, lc) and len(la) == 6
        assert a.timestep() == b.timestep() == c.timestep()            # the reductions left behind agree too


@pytest.mark.parametrize("integrator", ['euler', 'verlet'])
def test_other_integrators_vs_oracle(integrator):
    g, meta, pA = load_golden('dambreak20_cubic')
    P = O.Particles.from_aos(pA)
    w = O.wcsph(**{k: meta['consts'][k] for k in ('height', 'r0', 'rho0')}, useXSPH=True)
    with _ctx(meta, integrator=integrator) as ctx:
        ctx.upload(pA)
        for s in range(3):
            dt3, _ = O.step(P, w, 'cubic', integrator, True, False, 0.0, meta['fixed_h'])
            ctx.step(1, None, 0.0)
            cols = ctx.download_fields(list(STATE_FIELDS))
            for f in STATE_FIELDS:
                assert field_err(cols[f], getattr(P, f)) <= TOL, (s, f)
        assert np.allclose(ctx.dt_log()[-1], dt3, rtol=1e-10)


@pytest.mark.parametrize("N,kernel", [(60, 'wendland'), (150, 'cubic'), (150, 'gaussian'), (150, 'wendland')])
def test_dam_break_vs_oracle(N, kernel):
    """Larger seeded dam breaks: N=60 is the truncating regime (3h > cell), N=150 the fine-cell regime."""
    case = W.dam_break_case(N, seed=5)
    pA, c = case['pA'], case['consts']
    P = O.Particles.from_aos(pA)
    w = O.wcsph(c['height'], c['r0'], c['rho0'], True)
    cfg = capi.make_config(c, kernel, 'pec', capi.FP64, case['h'])
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        ctx.build_neighbours()
        grid, cells = ctx.cells()
        og = O.Grid(P)
        assert np.array_equal(cells, og.cell_ids())
        off, idx = ctx.neighbours_csr()
        ooff, oidx = og.neighbours_csr()
        assert np.array_equal(off, ooff)
        assert all(np.array_equal(a, b) for a, b in zip(_sorted_lists(off, idx), _sorted_lists(ooff, oidx)))
        for s in range(2):
            dt3, _ = O.step(P, w, kernel, 'pec', True, False, 0.05, case['h'])
            ctx.step(1, None, 0.05)
            cols = ctx.download_fields(list(STATE_FIELDS))
            for f in STATE_FIELDS:
                assert field_err(cols[f], getattr(P, f)) <= TOL, (s, f)
            assert np.allclose(ctx.dt_log()[-1], dt3, rtol=1e-10)


def test_reorder_cadence_does_not_change_results():
    """Physical re-sorting of the device state only permutes storage: fields agree to summation order."""
    case = W.dam_break_case(40, seed=6)
    outs = []
    for every in (1, 3, 1000):
        cfg = capi.make_config(case['consts'], 'wendland', 'pec', capi.FP64, case['h'], reorder_every=every)
        with capi.Context(cfg) as ctx:
            ctx.upload(case['pA'])
            ctx.step(7, None, 0.05)
            outs.append(ctx.download(case['pA'].copy()))
    for o in outs[1:]:
        for f in STATE_FIELDS:
            assert field_err(o[f], outs[0][f]) <= 1e-11, f


@pytest.mark.parametrize("precision", ['fp64', 'fp32'])
def test_counting_sort_gives_the_order_of_the_stable_radix_sort(precision, monkeypatch):
    """The default neighbour structure is a counting sort by cell whose in-cell order is made canonical (ascending
    storage slot) after the atomics; OSPH_SORT=radix is the stable LSD radix sort it replaced.  Same permutation, hence
    bit-identical results -- across a physical reorder of the state (3rd build) too."""
    case = W.dam_break_case(60, seed=12)
    prec = capi.FP64 if precision == 'fp64' else capi.FP32
    outs = []
    for mode in ('bin', 'radix', 'bin'):
        if mode == 'radix':
            monkeypatch.setenv("OSPH_SORT", "radix")
        else:
            monkeypatch.delenv("OSPH_SORT", raising=False)
        cfg = capi.make_config(case['consts'], 'cubic', 'pec', prec, case['h'])
        with capi.Context(cfg) as ctx:
            ctx.upload(case['pA'])
            ctx.step(6, None, 0.05)
            g, cells = ctx.cells()
            off, idx = ctx.neighbours_csr()
            outs.append((ctx.download(case['pA'].copy()).tobytes(), cells, off, idx, ctx.dt_log()))
            assert ctx.sync() == 0
    for o in outs[1:]:
        assert o[0] == outs[0][0]
        assert np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2]) and np.array_equal(o[3], outs[0][3])
        assert np.array_equal(o[4], outs[0][4])


@pytest.mark.parametrize("kernel,push,side,steps", [('cubic', 0.0, 100, 30), ('wendland', 20.0, 100, 30), ('gaussian', 0.0, 150, 12)])
def test_sort_cadence_reuses_the_binning_and_changes_nothing(kernel, push, side, steps, monkeypatch):
    """Between two sorts the device reuses the sorted order and the cell table (cells = pair radius + skin, re-sort as soon
    as pair radius + 2 x largest displacement exceeds the cell: k_grid_params).  Membership is decided on the current
    positions, so nothing may change: same neighbour sets as the oracle on the moved particles, fields equal to the run
    that sorts at every build (OSPH_SKIN=0) up to summation order, same dt.  push: fluid thrown at 20 m/s with a fixed time
    step -- a 10 % skin is used up every few steps and the run must re-sort on its own."""
    case = W.dam_break_case(side, seed=21)            # pair radius 0.8 m < reference cell 1 m: the fine-cell regime
    pA = case['pA'].copy()
    pA['vx'][pA['label'] == 0] += push
    runs = {}
    for skin in ('0', 'auto', '0.1'):
        monkeypatch.setenv("OSPH_SKIN", skin)
        cfg = capi.make_config(case['consts'], kernel, 'pec', capi.FP64, case['h'])
        with capi.Context(cfg) as ctx:
            ctx.upload(pA)
            ctx.step(steps, 4e-4 if push else None, 0.0)
            builds, sorts = ctx.sort_stats()
            off, idx = ctx.neighbours_csr()                   # queried on a binning that may be several steps old
            out = ctx.download(pA.copy())
            g, cells = ctx.cells()
            runs[skin] = dict(out=out, off=off, idx=idx, cells=cells, dts=ctx.dt_log(), builds=builds, sorts=sorts)
            assert ctx.sync() == 0
            if skin == '0.1':
                # neighbour sets of the final positions against the reference search of the oracle
                P = O.Particles.from_aos(out[~out['deleted']])
                ooff, oidx = O.Grid(P).neighbours_csr()
                assert np.array_equal(off, ooff)
                assert all(np.array_equal(a, b) for a, b in zip(_sorted_lists(off, idx), _sorted_lists(ooff, oidx)))
                assert np.array_equal(cells, O.Grid(P).cell_ids())
    ref = runs['0']
    assert ref['sorts'] == ref['builds'] >= steps
    for skin in ('auto', '0.1'):
        r = runs[skin]
        assert r['builds'] == ref['builds'] and 2 <= r['sorts'] < r['builds'], (skin, r['builds'], r['sorts'])
        assert np.array_equal(r['off'], ref['off']) and np.array_equal(r['idx'], ref['idx']) and np.array_equal(r['cells'], ref['cells'])
        assert np.allclose(r['dts'], ref['dts'], rtol=1e-12, atol=0)
        for f in STATE_FIELDS:
            assert field_err(r['out'][f], ref['out'][f]) <= 1e-11, (skin, f)
    if push:
        assert runs['0.1']['sorts'] > 3, "the thrown fluid must outrun a 10 % skin several times in 30 steps"


@pytest.mark.parametrize("bits", ['1', '3'])
def test_scalar_kernels_fused_into_their_producers_change_nothing(bits, monkeypatch):
    """OSPH_FUSE_SCALARS (off by default, measured not to pay): k_grid_params run by the last CTA of the predictor pass,
    k_timestep by the last CTA of the pair kernel.  Same operations on the same scalars: bit-identical state and dt."""
    case = W.dam_break_case(100, seed=4)
    outs = []
    for b in ('0', bits):
        monkeypatch.setenv("OSPH_FUSE_SCALARS", b)
        cfg = capi.make_config(case['consts'], 'cubic', 'pec', capi.FP64, case['h'])
        with capi.Context(cfg) as ctx:
            ctx.upload(case['pA'])
            ctx.step(9, None, 0.05)
            ctx.step(1, None, 0.05)
            outs.append((ctx.download(case['pA'].copy()).tobytes(), ctx.dt_log(), ctx.sort_stats()))
            assert ctx.sync() == 0
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


@pytest.mark.parametrize("kernel,precision", [('cubic', 'fp64'), ('wendland', 'fp64'), ('cubic', 'fp32'), ('wendland', 'fp32')])
def test_uniform_h_instantiation_gives_the_bits_of_the_general_one(kernel, precision, monkeypatch):
    """Solver(h=value): k_pair<.., UH> takes h_ij, 1/h_ij, the support test, h~ and the kernel normalisation of a fluid-fluid
    pair from constants that k_uh_constants formed with the operations of the general body; wall and gate neighbours (h = 0)
    go through the general bodies.  Same operations on the same operands: the state after several steps must be the SAME
    BITS as with OSPH_UH=0 (general instantiation), dt included, and the h-uniformity check of k_gather must stay silent.
    A set-up with a hand-edited fluid h under OSPH_H_KEEP must not use the instantiation."""
    case = W.dam_break_case(100, seed=9)
    prec = capi.FP64 if precision == 'fp64' else capi.FP32
    outs = []
    for env in ('0', '1'):
        monkeypatch.setenv("OSPH_UH", env)
        cfg = capi.make_config(case['consts'], kernel, 'pec', prec, case['h'])
        with capi.Context(cfg) as ctx:
            ctx.upload(case['pA'])
            ctx.compute()                                   # explicit-call path: one force evaluation of the uploaded state
            first = ctx.download(case['pA'].copy())
            ctx.step(6, None, 0.05)
            outs.append((ctx.download(case['pA'].copy()), ctx.dt_log(), ctx.pair_kernel_info(), first))
            assert ctx.sync() == 0
    (a, dta, ia, a0), (b, dtb, ib, b0) = outs
    assert ia[:2] == (7, 0)
    if not ib[2] & 1:
        pytest.skip("library built without PAIR_UH: the general instantiation ran both times")
    assert ib[:2] == (7, 7)
    if ib[2] & 4:
        # software-pipelined flush loop: the rare pairs of a flush (wall and gate neighbours) are summed after its common ones
        # -- equal to summation order, and bit-identical wherever a particle has no such neighbour
        # (after a step the global dt carries the difference to every particle: the bitwise part looks at the first evaluation)
        same = np.ones(len(a), bool)
        for f in STATE_FIELDS:
            assert field_err(b[f], a[f]) <= (1e-13 if precision == 'fp64' else 2e-6), f
            assert field_err(b0[f], a0[f]) <= (1e-14 if precision == 'fp64' else 1e-6), f
            same &= a0[f] == b0[f]
        assert same.mean() > 0.8, same.mean()
        assert np.allclose(dta, dtb, rtol=1e-12 if precision == 'fp64' else 1e-5, atol=0)
    else:
        for f in STATE_FIELDS:
            assert np.array_equal(a[f], b[f]) and np.array_equal(a0[f], b0[f]), f       # -0 == +0: the sign-bit clamps may differ there
        assert np.array_equal(dta, dtb)
    # smoothing length kept as uploaded: never the uniform-h instantiation, whatever the values are
    cfg = capi.make_config(case['consts'], kernel, 'pec', prec, case['h'], keep_h=True)
    with capi.Context(cfg) as ctx:
        ctx.upload(case['pA'])
        ctx.step(2, None, 0.05)
        assert ctx.pair_kernel_info()[:2] == (2, 0)


def test_uniform_h_instantiation_is_not_used_where_its_float_stand_in_could_overflow():
    """Float Wendland with a smoothing length below 1e-3: the pipelined uniform-h loop would evaluate its stand-in entries at
    q = 1 / h, beyond the float range of the degree-8 polynomial for h < 1.5e-5 -- the host selects the general instantiation
    there (its stand-in is q = 1).  Double, and float with the cubic spline, keep the uniform-h instantiation."""
    case = W.dam_break_case(20, seed=3)
    scale = 4e-4                                           # the same particle set shrunk to a spacing of 0.5 mm, h = 0.8 mm
    pA = case['pA'].copy()
    pA['x'] *= scale; pA['y'] *= scale
    consts = dict(case['consts'], r0=case['consts']['r0'] * scale)
    for kernel, prec, expect_uh in (('wendland', capi.FP32, False), ('wendland', capi.FP64, True), ('cubic', capi.FP32, True)):
        cfg = capi.make_config(consts, kernel, 'pec', prec, case['h'] * scale)
        with capi.Context(cfg) as ctx:
            ctx.upload(pA)
            ctx.step(2, None, 0.05)
            launches, uniform, flags = ctx.pair_kernel_info()
            cols = ctx.download_fields(['ax', 'ay', 'drho', 'x'])
            assert ctx.sync() & capi.S_NONFINITE == 0
        if flags & 1:
            assert (uniform == launches) == expect_uh, (kernel, prec, launches, uniform)
        assert all(np.all(np.isfinite(c)) for c in cols.values())


def test_fp32_mode_close_to_fp64():
    """Performance mode: float pair arithmetic on anchor-relative positions; drift bounded and reported."""
    case = W.dam_break_case(100, seed=7)
    res = {}
    for prec in (capi.FP64, capi.FP32):
        cfg = capi.make_config(case['consts'], 'wendland', 'pec', prec, case['h'])
        with capi.Context(cfg) as ctx:
            ctx.upload(case['pA'])
            ctx.step(20, None, 0.05)
            res[prec] = ctx.download(case['pA'].copy())
    a, b = res[capi.FP64], res[capi.FP32]
    r0 = case['r0']
    assert np.max(np.hypot(a['x'] - b['x'], a['y'] - b['y'])) < 1e-4 * r0      # position drift after 20 steps
    assert field_err(b['rho'], a['rho']) < 1e-6
    assert field_err(b['ax'], a['ax']) < 5e-3 and field_err(b['ay'], a['ay']) < 5e-3


def test_deleted_rows_untouched_and_index_space():
    """Deleted rows never reach the kernels and come back verbatim; indices are in compacted active order."""
    g, meta, pA = load_golden('dambreak20_wendland')
    pB = pA.copy()
    dead = np.flatnonzero(pB['label'] == 2)          # remove the temporary gate, as Solver.run() does after settling
    pB['deleted'][dead] = True
    pB['p'][dead] = -1e15
    act = pB[~pB['deleted']]
    P = O.Particles.from_aos(act)
    og = O.Grid(P)
    with _ctx(meta, keep_h=True) as ctx:
        ctx.upload(pB)
        assert ctx.num_active == len(act) and ctx.num_fluid == int((act['label'] == 0).sum())
        ctx.compute()
        off, idx = ctx.neighbours_csr()
        ooff, oidx = og.neighbours_csr()
        assert np.array_equal(off, ooff)
        assert all(np.array_equal(a, b) for a, b in zip(_sorted_lists(off, idx), _sorted_lists(ooff, oidx)))
        out = ctx.download(pB.copy())
        assert out[dead].tobytes() == pB[dead].tobytes()
        w = O.wcsph(**{k: meta['consts'][k] for k in ('height', 'r0', 'rho0')}, useXSPH=True)
        O.loop(P, w, og, 'wendland')
        for f in LOOP_FIELDS:
            assert field_err(out[f][~pB['deleted']], getattr(P, f)) <= TOL, f


def test_set_active_equals_download_mark_upload():
    """osph_set_active switches rows off (and on) on the device; the result is that of the round trip it replaces
    (download, mark deleted, upload: reference src/Solver.py:428-442), bit for bit, and a removed row keeps its last state
    in the record that comes back."""
    g, meta, pA = load_golden('dambreak20_wendland')
    gate = pA['label'] == 2
    with _ctx(meta, keep_h=True) as a, _ctx(meta, keep_h=True) as b:
        a.upload(pA); b.upload(pA)
        a.step(3, None, 0.05); b.step(3, None, 0.05)
        # old way: the whole array crosses PCIe twice
        host = a.download(pA.copy())
        host['deleted'][gate] = True
        a.upload(host)
        # new way: one byte per row
        b.set_active(~gate)
        assert b.num_active == a.num_active == int((~gate).sum()) and b.num_fluid == a.num_fluid
        a.step(3, None, 0.0); b.step(3, None, 0.0)
        ra, rb = a.download(host.copy()), b.download(pA.copy())
        assert rb['deleted'][gate].all() and not rb['deleted'][~gate].any()
        assert ra.tobytes() == rb.tobytes()
        assert np.array_equal(a.dt_log()[-3:], b.dt_log()[-3:])
        # rows come back: they resume from their records
        b.set_active(np.ones(len(pA), dtype=bool))
        assert b.num_active == len(pA)
        back = b.download(pA.copy())
        assert not back['deleted'].any()
        for f in STATE_FIELDS:
            assert np.array_equal(back[f][gate], host[f][gate]), f
        with pytest.raises(capi.OsphError):
            b.set_active(np.ones(len(pA) - 1, dtype=bool))


def test_near_pos_matches_oracle():
    g, meta, pA = load_golden('tank24_wendland_coupled')
    P = O.Particles.from_aos(pA)
    og = O.Grid(P)
    rng = np.random.default_rng(1)
    with _ctx(meta, keep_h=True) as ctx:
        ctx.upload(pA)
        ctx.build_neighbours()
        for _ in range(20):
            x, y = rng.uniform(0, 1, 2)
            h = float(rng.uniform(0.03, 0.12))
            hh, q, r, idx = ctx.near_pos(x, y, h)
            oh, oq, orr, oidx = og.near_pos(x, y, h)
            order = np.argsort(oidx)
            assert np.array_equal(idx, oidx[order])
            assert np.array_equal(r, orr[order]) and np.array_equal(q, oq[order]) and np.array_equal(hh, oh[order])


def test_kinetic_energy_and_timestep_edges():
    g, meta, pA = load_golden('dambreak20_cubic')
    P = O.Particles.from_aos(pA)
    with _ctx(meta) as ctx:
        ctx.upload(pA)
        assert ctx.kinetic_energy() == pytest.approx(O.kinetic_energy(P, P.fluid), rel=1e-13)
        assert ctx.timestep() == O.timestep(P, P.fluid)          # strict-IEEE reduction: bit-exact
    # no fluid particle: the reference raises from np.min([]); the ABI returns OSPH_E_NO_FLUID
    walls = pA[pA['label'] != 0]
    with _ctx(meta) as ctx:
        ctx.upload(walls)
        with pytest.raises(capi.OsphError) as e:
            ctx.timestep()
        assert e.value.code == -5
    # nothing uploaded
    with _ctx(meta) as ctx:
        with pytest.raises(capi.OsphError):
            ctx.compute()


def test_async_export_snapshots_in_stream_order():
    """osph_export_begin/_end: the columns are those of the moment of the begin, two tickets in flight, the copy
    overlaps later steps (include/osph.h; replaces the blocking copies of src/Solver.py:477-486)."""
    g, meta, pA = load_golden('dambreak20_cubic')
    names = ['x', 'y', 'p', 'c']
    with _ctx(meta) as ctx:
        ctx.upload(pA)
        ctx.step(2)
        want0 = ctx.download_fields(names)
        t0 = ctx.export_begin(names)
        ctx.step(3)                                   # state moves on while ticket 0 is in flight
        want1 = ctx.download_fields(names[:2])
        t1 = ctx.export_begin(names[:2])
        with pytest.raises(capi.OsphError):           # ring of two
            ctx.export_begin(names)
        ctx.step(1)
        got0 = ctx.export_end(t0)
        got1 = ctx.export_end(t1)
        with pytest.raises(capi.OsphError):           # a ticket can be ended once
            ctx.export_end(t0)
        for f in names:
            assert np.array_equal(got0[f], want0[f]), f
        for f in names[:2]:
            assert np.array_equal(got1[f], want1[f]), f
        assert not np.array_equal(want0['x'], want1['x'])
        t2 = ctx.export_begin(['rho'])                # slots are reusable, with a different column count
        assert np.array_equal(ctx.export_end(t2)['rho'], ctx.download_fields(['rho'])['rho'])
    # row space: one value per uploaded row, deleted rows keep what the host uploaded (what Solver._store appends)
    pB = pA.copy()
    dead = np.zeros(len(pB), dtype=bool)
    dead[5::7] = True
    pB['deleted'] = dead
    pB['p'][dead] = -1e15
    with _ctx(meta) as ctx:
        ctx.upload(pB)
        ctx.step(2)
        act = ctx.download_fields(['x', 'p', 'c'])
        rows = ctx.export_end(ctx.export_begin(['x', 'p', 'c'], rows=True))
        for f in ('x', 'p', 'c'):
            assert len(rows[f]) == len(pB)
            assert np.array_equal(rows[f][~dead], act[f]), f
            assert np.array_equal(rows[f][dead], pB[f][dead]), f


def test_row_transfers():
    """osph_download_rows / osph_upload_rows: a few host rows by number, across a physical reorder of the device
    storage, with deleted rows in between (src/Solver.py:381-398 moves the Coupled rows this way)."""
    g, meta, pA = load_golden('tank24_wendland_coupled')
    pB = pA.copy()
    dead = np.zeros(len(pB), dtype=bool)
    dead[3::11] = True
    pB['deleted'] = dead
    live = np.flatnonzero(~dead)
    rows = live[::5][:40].astype(np.int64)[::-1].copy()            # unordered on purpose
    with _ctx(meta) as ctx:
        ctx.upload(pB)
        ctx.step(3)                                                # build #2 reorders the storage physically
        full = ctx.download(pB.copy())
        part = pB.copy()
        ctx.download_rows(rows, part)
        assert part[rows].tobytes() == full[rows].tobytes()
        untouched = np.setdiff1d(np.arange(len(pB)), rows)
        assert part[untouched].tobytes() == pB[untouched].tobytes()
        # write the rows back shifted, everything else must stay as it was
        edit = full.copy()
        edit['y'][rows] += 0.125
        edit['vx'][rows] = -3.0
        edit['c'][rows] = 7.0
        ctx.upload_rows(rows, edit)
        after = ctx.download(pB.copy())
        assert after[live].tobytes() == edit[live].tobytes()
        with pytest.raises(capi.OsphError):                        # deleted rows have no device state
            ctx.download_rows(np.array([np.flatnonzero(dead)[0]], dtype=np.int64), part)
        with pytest.raises(capi.OsphError):
            ctx.download_rows(np.array([len(pB)], dtype=np.int64), part)
        ctx.step(1)                                                # the neighbour structure is rebuilt after the upload
        assert np.all(np.isfinite(ctx.download(pB.copy())['ax'][live]))


def test_upload_fields_roundtrip_and_single_particle():
    g, meta, pA = load_golden('block20_cubic_nobnd')
    with _ctx(meta) as ctx:
        ctx.upload(pA)
        cols = ctx.download_fields(['x', 'vy', 'c'])
        assert np.array_equal(cols['x'], pA['x']) and np.array_equal(cols['c'], pA['c'])
        ctx.upload_fields({'vy': cols['vy'] * 2.0})
        assert np.array_equal(ctx.download_fields(['vy'])['vy'], pA['vy'] * 2.0)
    one = pA[:1].copy()
    with _ctx(meta) as ctx:                                       # a lone particle only sees itself
        ctx.upload(one)
        ctx.compute()
        off, idx = ctx.neighbours_csr()
        assert list(off) == [0, 1] and list(idx) == [0]
        out = ctx.download(one.copy())
        assert out['ax'][0] == 0.0 and out['ay'][0] == -9.81 and out['drho'][0] == 0.0


def test_particle_on_upper_grid_edge():
    """Extent an exact multiple of the cell: the reference bins x == xmax into column ncx, which
    wraps into the next row (or past the table).  The CUDA path reproduces the wrap bit for bit."""
    c = W.wcsph_constants(2.0, 0.25, 1000.0, True)
    xs, ys = np.meshgrid(np.linspace(0.0, 2.0, 9), np.linspace(0.0, 3.0, 13), indexing='ij')
    pA = np.zeros(xs.size, dtype=O.particle_dtype)
    pA['x'] = xs.ravel(); pA['y'] = ys.ravel() + 0.01 * np.sin(7 * xs.ravel())
    pA['y'][0] = 0.0; pA['y'][-1] = 3.0
    pA['m'] = 62.5; pA['rho'] = 1000.0; pA['h'] = 0.5
    pA['label'][:13] = 1; pA['h'][:13] = 0.0
    P = O.Particles.from_aos(pA)
    og = O.Grid(P)
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, 0.5, keep_h=True)
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        ctx.build_neighbours()
        grid, cells = ctx.cells()
        assert grid['cell_size'] == 1.0 and (grid['ncx'], grid['ncy']) == (og.params['ncx'], og.params['ncy'])
        assert np.array_equal(cells, og.cell_ids())
        off, idx = ctx.neighbours_csr()
        ooff, oidx = og.neighbours_csr()
        assert np.array_equal(off, ooff)
        assert all(np.array_equal(a, b) for a, b in zip(_sorted_lists(off, idx), _sorted_lists(ooff, oidx)))
        assert (ctx.sync() & capi.S_UNBINNED) == (capi.S_UNBINNED if og.rc != 0 else 0)
