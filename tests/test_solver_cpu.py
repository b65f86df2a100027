"""CPU: the HOST logic of the re-hosted Solver (offshore-sph_b200/src/Solver.py) with the oracle standing in for the
device context (tests/oracle_context.py): same control flow as reference src/Solver.py:366-465, checked against the
reference Solver's own end state (golden `solver_dambreak12_wendland`).  The GPU twin of these tests is
tests/test_gpu_solver.py, which runs the same Solver on the real context."""
import numpy as np
import pytest

from conftest import field_err, load_golden
from oracle import oracle as O
from oracle_context import OracleContext
from osph_b200 import capi, workloads as W


@pytest.fixture
def oracle_backend(monkeypatch):
    monkeypatch.setenv("OSPH_QUIET", "1")
    monkeypatch.setattr(capi, "Context", OracleContext)
    monkeypatch.setattr(capi, "default_device", lambda: 0)
    OracleContext.instances.clear()
    yield OracleContext
    OracleContext.instances.clear()


def _dam_break_solver(meta, **kw):
    from src.Solver import Solver
    from src.Methods.WCSPH import WCSPH
    from src.Kernels.Wendland import Wendland
    from src.Integrators.PEC import PEC
    r0, pA = W.dam_break(meta['N'])
    method = WCSPH(height=25.0, r0=r0, rho0=1000.0, useXSPH=True, Pb=0, useSummationDensity=False)
    s = Solver(method, PEC(useXSPH=True, strict=False), Wendland(), meta['duration'], incrementalWriteout=False,
               h=meta['hfac'] * r0, maxSettle=meta['maxSettle'], **kw)
    s.addParticles(pA)
    return s


@pytest.mark.parametrize("sync_export", [False, True])
def test_solver_control_flow_matches_reference_solver(oracle_backend, monkeypatch, sync_export):
    """Settling -> temp-boundary removal -> time stepping; asynchronous (default) and blocking export."""
    if sync_export:
        monkeypatch.setenv("OSPH_SYNC_EXPORT", "1")
    else:
        monkeypatch.delenv("OSPH_SYNC_EXPORT", raising=False)
    g, meta, _ = load_golden('solver_dambreak12_wendland')
    s = _dam_break_solver(meta)
    s.setup()
    s.run()
    ctx = oracle_backend.instances[-1]
    ref = np.frombuffer(g['final'].tobytes(), dtype=O.particle_dtype)
    assert s.t_step == int(g['t_step'])
    assert np.allclose(s.dt_a, g['dt_a'], rtol=1e-10, atol=0)
    assert np.allclose(s.dt_c, g['dt_c'], rtol=1e-10, atol=0) and np.allclose(s.dt_f, g['dt_f'], rtol=1e-10, atol=0)
    assert s.settleTime == pytest.approx(float(g['settleTime']), rel=1e-10)
    assert s.t == pytest.approx(float(g['t']), rel=1e-10)
    assert np.array_equal(s.particleArray['deleted'], ref['deleted'])
    act = ~ref['deleted']
    for f in ('x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay', 'drho', 'h', 'c'):
        assert field_err(s.particleArray[f][act], ref[f][act]) <= 1e-9, f
    assert np.all(s.particleArray['p'][~act] == -1e15)
    # export: one full-length array per step, last one = the reference's, nothing left in flight
    assert len(s.export['x']) == int(g['n_export']) == s.t_step
    assert all(len(a) == len(s.particleArray) for a in s.export['x'])
    assert field_err(s.export['x'][-1], g['export_x_last']) <= 1e-9
    assert not s._export_pending and not ctx._tickets
    if sync_export:
        assert ctx.calls.get('download_fields', 0) == s.t_step and 'export_begin' not in ctx.calls
    else:
        assert ctx.calls.get('export_begin', 0) == ctx.calls.get('export_end', 0) == s.t_step
    # the whole array crosses the boundary only at setup and at the end of the run; the gate is removed on the device
    # (osph_set_active: the gate's own rows and a one-byte-per-row mask)
    assert ctx.calls['upload'] == 1 and ctx.calls['download'] <= 3 and ctx.calls['set_active'] == 1


def test_solver_rejects_empty_and_inconsistent_particle_sets(oracle_backend):
    """reference src/Solver.py:172-174, 352-354: `raise Exception('No or invalid particles set!')`."""
    g, meta, _ = load_golden('solver_dambreak12_wendland')
    from src.Solver import Solver
    s = _dam_break_solver(meta)
    s.particleArray = None
    with pytest.raises(Exception, match='No or invalid particles'):
        s.setup()
    s = _dam_break_solver(meta)
    s.setup()
    s.particleArray['deleted'][0] = True                # active count no longer matches what setup() saw
    with pytest.raises(Exception, match='No or invalid particles'):
        s.run()
    assert isinstance(s, Solver)


def test_coupling_rows_only_moves_only_the_coupled_rows(oracle_backend):
    """couplingRowsOnly=True: the whole-array round trips around the coupled rows (reference src/Solver.py:381-398)
    become row transfers; both modes give the same simulation."""
    from src.Solver import Solver
    from src.Methods.WCSPH import WCSPH
    from src.Kernels.Wendland import Wendland
    from src.Integrators.PEC import PEC
    from src.Integrators.NewmarkBeta import NewmarkBeta
    from src.Common import ParticleType
    runs = []
    for rows_only in (False, True):
        case = W.tank_case(10, h=1.3 / 10, useXSPH=True, seed=2, coupled_row=True, perturbed=False)
        pA = case['pA']
        c = pA['label'] == ParticleType.Coupled
        nc = int(c.sum())
        y0 = pA['y'][c].copy()
        seen = []

        def coupling(arr, solver, c=c, y0=y0, seen=seen):
            a = solver.couplingIntegrator.acceleration(solver.dt, np.full(c.sum(), -50.0), arr['y'][c] - y0, arr['vy'][c])
            arr['ay'][c] = a
            seen.append(float(a[0]))
            return arr

        nb = NewmarkBeta(0.25, 0.5, np.eye(nc) * 50.0, np.eye(nc) * 1e4, np.eye(nc) * 10.0)
        method = WCSPH(height=1.0, r0=case['r0'], rho0=1000.0, useXSPH=True, Pb=0)
        s = Solver(method, PEC(useXSPH=True, strict=False), Wendland(), 0.001, incrementalWriteout=False, h=1.3 / 10,
                   maxSettle=2, coupling=coupling, couplingIntegrator=nb, couplingProperties={}, exportProperties=['y'],
                   couplingRowsOnly=rows_only)
        s.addParticles(pA)
        s.setup()
        s.run()
        ctx = oracle_backend.instances[-1]
        assert len(seen) == s.t_step
        if rows_only:
            gate = 1 if (pA['label'] == 2).any() else 0          # the removal of a temporary boundary moves its rows once more
            assert ctx.calls['download_rows'] == ctx.calls['upload_rows'] == 3 * s.t_step + gate
            assert ctx.calls['upload'] == 1 and ctx.calls['set_active'] == 1     # setup; gate removal on the device
        else:
            assert 'download_rows' not in ctx.calls and ctx.calls['upload'] >= 3 * s.t_step
        runs.append(s)
    a, b = runs
    assert a.t_step == b.t_step and a.dt_a == b.dt_a
    for f in ('x', 'y', 'vx', 'vy', 'rho', 'ax', 'ay'):
        assert np.array_equal(a.particleArray[f], b.particleArray[f]), f
    assert np.any(a.particleArray['y'][a.particleArray['label'] == ParticleType.Coupled] != 0)


@pytest.mark.parametrize("example, args", [("containment", ['--nx', '12', '--duration', '0.002', '--max-settle', '3']),
                                            ("dam_break", ['--n', '10', '--duration', '0.002', '--max-settle', '3'])])
def test_shipped_examples_run_on_the_oracle_backend(oracle_backend, example, args, tmp_path):
    """The package's own example scripts (Containment: dynamic h, XSPH off, `Solver(h=None)`; both call timing() and
    save()) drive the Solver end to end on CPU."""
    import importlib.util
    import os
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("ex_" + example, os.path.join(PKG, "examples", example + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    s = mod.main(args + ['--out', str(tmp_path / "run")])
    pa = s.particleArray
    assert s.t >= 0.002 and s.t_step > 0 and np.all(np.isfinite(pa['x'])) and np.all(np.isfinite(pa['rho']))
    fluid = (pa['label'] == 0) & ~pa['deleted']
    assert np.all(pa['h'][fluid] > 0)
    for key in s.exportProperties:
        assert len(s.export[key]) == s.t_step
    assert any(f.startswith("run") for f in os.listdir(tmp_path))          # save(): .hdf5 with h5py, else .npz
