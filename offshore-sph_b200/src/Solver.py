"""Solver with the constructor, attributes and control flow of reference src/Solver.py:46-531, re-hosted so
that the particle state stays resident in HBM across the five per-step calls.

Per step (same order as reference src/Solver.py:366-465):
    dt = osph_timestep            <- TimeStep().compute(...)                       (:368-372)
    osph_predict                  <- integrator.predict(dt, pA[f_indexes], damping) (:380)
    osph_build_neighbours         <- nn.update(pA[indexes]) + h refresh             (:238-246)
    osph_compute                  <- _loop(pA[indexes], kernel..., method, nn)      (:254)
    [coupling callback on a downloaded host copy]                                   (:389-392)
    osph_correct                  <- integrator.correct(...)                        (:396)
    _store: only the exportProperties columns are downloaded, asynchronously        (:477-486)

`solver.particleArray` stays a plain numpy array; it is refreshed from the device when the run ends, when the
settling phase ends, around coupling callbacks and on save().  Backend knobs come from the environment so
that example scripts run unedited: OSPH_PRECISION=fp64|fp32, OSPH_DEVICE=<ordinal>, OSPH_STORE_FULL=1
(keep the reference's per-step full copies in `solver.data`), OSPH_SYNC_EXPORT=1 (blocking per-step export; by
default the exported columns of step k cross PCIe while step k+1 is computed and appear in `solver.export` one
step later -- always complete when run() returns or save() is called).
"""
import os
from time import perf_counter
from typing import List

import numpy as np

from src.Common import particle_dtype, ParticleType
from src.Tools.NNLinkedList import NNLinkedList
from src.Tools.SolverTools import findActive
from osph_b200 import capi

try:
    from tqdm import tqdm
except ImportError:          # progress bars are cosmetic
    tqdm = None


class _NoBar:
    def __init__(self, *a, **k): pass
    def update(self, *a): pass
    def close(self): pass


def _bar(**kw):
    if tqdm is None or os.environ.get("OSPH_QUIET"):
        return _NoBar()
    return tqdm(**kw)


def println(text: str):
    if not os.environ.get("OSPH_QUIET"):
        print('\n{0}'.format(text))


class Solver:
    def __init__(self, method, integrator, kernel, duration: float = 1.0, quick: bool = True,
                 incrementalWriteout: bool = True, incrementalFile: str = "export", incrementalFreq: int = 1000,
                 exportProperties: List[str] = ['x', 'y', 'p'], kE: float = 0.8, maxSettle: int = 500,
                 timeStep: float = None, h: float = None, coupling=None, couplingIntegrator=None,
                 couplingProperties=None, customSettle=None, damping: float = 0.05, couplingRowsOnly: bool = None):
        self.method = method
        self.integrator = integrator
        self.kernel = kernel
        self.nn = NNLinkedList(scale=2.0)

        self.duration = duration
        self.quick = quick
        self.kEF = kE
        self.maxSettle = maxSettle
        self.timeStep = timeStep
        self.h = h
        self.ts_error = 0
        self.damping = damping

        self.coupling = coupling
        self.couplingIntegrator = couplingIntegrator
        self.couplingProperties = couplingProperties
        # True: only the Coupled rows travel between host and device around couplingIntegrator.predict / .correct and
        # the coupling callback (osph_download_rows / osph_upload_rows); the callback then sees stale fluid rows in
        # the array it is handed and should query the fluid through solver.probe_pressure().  Default (reference
        # behaviour): the whole array is refreshed and re-uploaded.
        if couplingRowsOnly is None:
            couplingRowsOnly = bool(os.environ.get("OSPH_COUPLING_ROWS"))
        self.couplingRowsOnly = bool(couplingRowsOnly)
        self.customSettle = customSettle

        self.incrementalWriteout = incrementalWriteout
        self.incrementalFreq = incrementalFreq
        self.incrementalFile = incrementalFile
        self.exportProperties = list(exportProperties)

        self.particleArray = None
        self.num_particles = 0
        self.data = []
        self.export = {}
        self.settleTime = 0.0
        self.dt_a, self.dt_c, self.dt_f = [], [], []
        self.t = 0.0
        self.dt = 0.0
        self.t_step = 0
        self.settled = False
        self.timing_data = {k: 0.0 for k in ('total', 'storage', 'integrate_correction', 'integrate_prediction',
                                             'compute', 'time_step', 'neighbour_hood', 'coupling')}
        self._ctx = None
        self._host_dirty = False      # device state is newer than self.particleArray
        self._store_full = bool(os.environ.get("OSPH_STORE_FULL"))
        # export columns cross PCIe on a copy stream while the next step runs (osph_export_begin / _end);
        # OSPH_SYNC_EXPORT=1 restores the blocking per-step download
        self._async_export = not os.environ.get("OSPH_SYNC_EXPORT")
        self._export_pending = []

    # ------------------------------------------------------------------------------------------
    def load(self, file: str):
        """Continue from a previous export (reference :143-158): .hdf5 via h5py when present, else .npz."""
        if file.endswith('.npz') or not _have_h5py():
            with np.load(file if file.endswith('.npz') else file + '.npz', allow_pickle=False) as z:
                self.particleArray = z['particleArray'].view(particle_dtype).reshape(-1)
                self.dt_a, self.dt_c, self.dt_f = list(z['dt_a']), list(z['dt_c']), list(z['dt_f'])
        else:
            import h5py
            with h5py.File(file, 'r') as h5f:
                self.particleArray = h5f['particleArray'][:]
                self.dt_a, self.dt_c, self.dt_f = list(h5f['dt_a'][:]), list(h5f['dt_c'][:]), list(h5f['dt_f'][:])

    def addParticles(self, particles: np.array):
        if self.particleArray is None:
            self.particleArray = np.ascontiguousarray(particles)
        else:
            self.particleArray = np.concatenate((self.particleArray, particles))

    # ------------------------------------------------------------------------------------------
    def _integrator_name(self):
        name = getattr(self.integrator, 'osph_name', None)
        if name is None:
            raise TypeError('integrator %r has no device implementation' % type(self.integrator).__name__)
        return name

    def _make_context(self):
        kname = getattr(self.kernel, 'osph_name', None)
        if kname is None:
            raise TypeError('kernel %r has no device implementation' % type(self.kernel).__name__)
        cfg = capi.make_config(self.method.constants(), kname, self._integrator_name(), capi.precision_from_env(),
                               self.h, integrator_xsph=getattr(self.integrator, 'useXSPH', False),
                               strict=getattr(self.integrator, 'strict', False), device=capi.default_device())
        cfg.nn_scale = self.nn.scale
        return capi.Context(cfg)

    def _masks(self):
        self.num_particles, self.indexes = findActive(self.num_particles, self.particleArray)
        pa = self.particleArray
        self.f_indexes = (pa['label'] == ParticleType.Fluid) & (pa['deleted'] == False)      # noqa: E712
        self.c_indexes = (pa['label'] == ParticleType.Coupled) & (pa['deleted'] == False)    # noqa: E712
        self.fluid_count = int(np.sum(self.f_indexes))
        self._c_rows = np.flatnonzero(self.c_indexes).astype(np.int64)

    def _pull(self):
        """Refresh the host mirror from the device."""
        if self._ctx is not None and self._host_dirty:
            self._ctx.download(self.particleArray)
            self._host_dirty = False

    def _push(self):
        """Replace the device state by the host mirror (after user code edited it)."""
        self._ctx.upload(self.particleArray)
        self._host_dirty = False

    def setup(self):
        println('Starting setup.')
        if self.particleArray is None or len(self.particleArray) == 0:
            raise Exception('No or invalid particles set!')
        self.particleArray = np.ascontiguousarray(self.particleArray)
        self._masks()
        println('{0} total particles, {1} fluid particles.'.format(self.num_particles, self.fluid_count))
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = self._make_context()
        self.nn._bind(self._ctx)
        # h, hydrostatic density, p, c of the fluid rows: on the device (osph_initialize)
        self._ctx.upload(self.particleArray)
        self._ctx.initialize()
        self._host_dirty = True
        self._pull()
        self.data.append(self.particleArray[:])
        self.kE = self._ctx.kinetic_energy() * self.kEF
        for key in self.exportProperties:
            self.export[key] = []
        println('Setup complete.')

    # ------------------------------------------------------------------------------------------
    def _minTimeStep(self) -> float:
        start = perf_counter()
        m, c, f = self._ctx.timestep()
        if m < 1e-6:
            self.ts_error = len(self.dt_a)
        self.dt_c.append(c); self.dt_f.append(f); self.dt_a.append(m)
        self.timing_data['time_step'] += perf_counter() - start
        return m

    def _compute(self):
        start = perf_counter()
        self._ctx.build_neighbours()
        self.timing_data['neighbour_hood'] += perf_counter() - start
        start = perf_counter()
        self._ctx.compute()
        self._host_dirty = True
        self.timing_data['compute'] += perf_counter() - start

    def _host_rows(self, mask, fn, *args):
        if self.couplingRowsOnly and mask is self.c_indexes:
            rows = self._c_rows
            self._ctx.download_rows(rows, self.particleArray)
            self.particleArray[rows] = fn(*args[:1], self.particleArray[rows], *args[1:])
            self._ctx.upload_rows(rows, self.particleArray)
            return
        self._pull()
        self.particleArray[mask] = fn(*args[:1], self.particleArray[mask], *args[1:])
        self._push()

    def run(self):
        start_all = perf_counter()
        if self.particleArray is None or len(self.particleArray) == 0 or \
                int(np.sum(~self.particleArray['deleted'])) != self.num_particles:
            raise Exception('No or invalid particles set!')
        println('Started solving...')
        ctx = self._ctx
        t_step = 0
        self.t = 0.0
        println('Settling particles...')
        self.settled = False
        sbar = _bar(total=self.maxSettle, desc='Settling', leave=False)
        tbar = _NoBar()

        # OSPH_MAX_STEPS=<n>: stop after n steps whatever the duration (smoke runs of unedited example scripts)
        max_steps = int(os.environ.get("OSPH_MAX_STEPS", "0")) or None
        while self.t < self.duration and (max_steps is None or t_step < max_steps):
            if self.timeStep is None:
                self.dt = self._minTimeStep()
            else:
                self.dt_a.append(self.timeStep)
                self.dt = self.timeStep

            if self.integrator.isMultiStage():
                self._compute()

            start = perf_counter()
            ctx.predict(self.dt, self.damping)
            self._host_dirty = True
            if self.coupling is not None:
                self._host_rows(self.c_indexes, self.couplingIntegrator.predict, self.dt, self.damping)
            self.timing_data['integrate_prediction'] += perf_counter() - start

            self._compute()

            if self.coupling is not None:
                start = perf_counter()
                if self.couplingRowsOnly:
                    ctx.download_rows(self._c_rows, self.particleArray)
                    self.particleArray = self.coupling(self.particleArray, self)
                    ctx.upload_rows(self._c_rows, self.particleArray)
                else:
                    self._pull()
                    self.particleArray = self.coupling(self.particleArray, self)
                    self._push()
                self.timing_data['coupling'] += perf_counter() - start

            start = perf_counter()
            ctx.correct(self.dt, self.damping)
            self._host_dirty = True
            if self.coupling is not None:
                self._host_rows(self.c_indexes, self.couplingIntegrator.correct, self.dt, self.damping)
            self.timing_data['integrate_correction'] += perf_counter() - start

            start = perf_counter()
            self._store(t_step)
            self.timing_data['storage'] += perf_counter() - start

            if self.settled:
                self.t += self.dt
            t_step += 1

            # Status of the device every 32 steps (the loop synchronises once per step anyway, for dt): a domain that outgrew
            # its cell table is re-sized at the next build instead of running on coarsened cells for the rest of the run,
            # and a non-finite state is reported when it appears, not after the last step.
            if t_step % 32 == 0:
                self._poll_status()

            if not self.settled and t_step > 1:
                ke = 1e12; cs = False
                if self.customSettle is None:
                    ke = ctx.kinetic_energy()
                else:
                    self._pull()
                    cs = self.customSettle(self.particleArray, self)
                if (ke < self.kE) or (t_step > self.maxSettle) or (cs == True):     # noqa: E712
                    if t_step > self.maxSettle:
                        println('WARNING! Maximum settle steps reached, check configuration, maybe increase spacing '
                                'between wall and particles.')
                    sbar.close()
                    # remove the temporary boundary (reference :428-442) ON the device: only the gate's own rows and a
                    # one-byte-per-row mask cross PCIe (osph_set_active); labels and deleted flags never change on the
                    # device, so the host mirror's copies are current even while its state columns are stale
                    self._export_drain()
                    pa = self.particleArray
                    inds = np.flatnonzero((pa['label'] == ParticleType.TempBoundary) & ~pa['deleted']).astype(np.int64)
                    if len(inds):
                        ctx.download_rows(inds, pa)
                        pa['p'][inds] = -1e15
                        ctx.upload_rows(inds, pa)
                    pa['deleted'][inds] = True
                    self._masks()
                    ctx.set_active(~pa['deleted'])
                    self._host_dirty = True
                    self.settleTime = sum(self.dt_a)
                    self.damping = 0.0
                    self.settled = True
                    println('Settling Complete.')
                    tbar = _bar(total=self.duration, desc='Time-stepping', unit='s', leave=False)
                else:
                    sbar.update(1)

            if len(self.data) > self.incrementalFreq:
                self.data.pop(0)
            if self.settled:
                tbar.update(self.dt)

        tbar.close()
        self._export_drain()
        self._pull()
        self._poll_status()
        self.timing_data['total'] = perf_counter() - start_all
        self.t_step = t_step
        println('Solved!')
        println('Solved {0} particles for {1:f} [s].'.format(self.num_particles, self.duration))
        println('Completed solve in {0:f} [s] and {1} steps'.format(self.timing_data['total'], t_step))

    def probe_pressure(self, x, y, h):
        """SPH-interpolated (rho, p) at the points (x[k], y[k]) in ONE device launch: the batched form of the
        per-node `pressure_SPH` helper of examples/IceBreak.py:252-285 for coupling callbacks."""
        return self._ctx.probe_pressure(x, y, h)

    def _store(self, t_step: int):
        if self._store_full:
            self._pull()
            self.data.append(np.copy(self.particleArray))
        if self.exportProperties:
            if self._async_export:
                # step k's columns travel while step k+1 is computed; step k-1's are merged into the lists now
                self._export_pending.append(self._ctx.export_begin(self.exportProperties, rows=True))
                while len(self._export_pending) > 1:
                    self._export_finish()
            else:
                self._export_merge(self._ctx.download_fields(self.exportProperties), self.indexes)
        if self.incrementalWriteout and t_step % self.incrementalFreq == 0:
            self.save('{0}-{1}.hdf5'.format(self.incrementalFile, t_step), printLocation=False)

    def _export_merge(self, cols, indexes):
        for key in cols:
            full = np.copy(self.particleArray[key])
            full[indexes] = cols[key]
            self.export[key].append(full)

    def _export_finish(self):
        # row-space columns: already the full per-row arrays (deleted rows keep their uploaded values)
        for key, col in self._ctx.export_end(self._export_pending.pop(0)).items():
            self.export[key].append(col)

    def _poll_status(self):
        """osph_sync: reads and clears the device's status bits; a coarsened grid makes the next build re-size its table."""
        status = self._ctx.sync()
        self.device_status = getattr(self, 'device_status', 0) | status
        if status & capi.S_NONFINITE and not getattr(self, '_warned_nonfinite', False):
            self._warned_nonfinite = True
            println('WARNING! non-finite particle state produced on the device.')
        if status & capi.S_H_NOT_UNIFORM:
            # the uniform-h pair kernel took h of every fluid particle from Solver(h=...), and one carried another value:
            # the rates of its neighbours are wrong -- never continue silently (OSPH_UH=0 selects the general kernel)
            raise RuntimeError('device status OSPH_S_H_NOT_UNIFORM: a fluid particle reached the pair kernel with a '
                               'smoothing length other than Solver(h=...); rerun with OSPH_UH=0 and report this')
        return status

    def _export_drain(self):
        while self._export_pending:
            self._export_finish()

    # ------------------------------------------------------------------------------------------
    def timing(self):
        total = self.timing_data['total'] or 1.0
        rows = [(k, round(v, 3), round(v / total * 100, 2)) for k, v in self.timing_data.items()]
        println('Detailed timing statistics:')
        w = max(len(r[0]) for r in rows)
        println('\n'.join(['%-*s  %10s  %14s' % (w, 'Name', 'Time [s]', 'Percentage [%]')] +
                          ['%-*s  %10.3f  %14.2f' % (w, *r) for r in rows]))
        if self._ctx is not None:
            dev = self._ctx.timers()
            println('Device time per phase [s]: ' + ', '.join('%s %.3f' % kv for kv in dev.items()))

    def save(self, location: str, printLocation: bool = True, extraProperties: dict = None):
        """gzip HDF5 like the reference (:497-531) when h5py is importable, else the same datasets as .npz."""
        self._export_drain()
        println('Starting file export.')
        self._pull()
        datasets = dict(particleArray=self.particleArray, dt_a=np.asarray(self.dt_a), dt_c=np.asarray(self.dt_c),
                        dt_f=np.asarray(self.dt_f), settleTime=np.asarray(self.settleTime))
        for key in self.exportProperties:
            if self.export.get(key):
                datasets[key] = np.stack(self.export[key])
        if extraProperties:
            for key, value in extraProperties.items():
                try:
                    datasets[key] = np.asarray(value)
                except Exception:
                    pass
        if _have_h5py():
            import h5py
            with h5py.File(location, 'w') as h5f:
                for key, value in datasets.items():
                    if np.ndim(value) == 0:
                        h5f.create_dataset(key, data=value)
                    else:
                        h5f.create_dataset(key, data=value, shuffle=True, compression="gzip")
        else:
            location = location + '.npz'
            datasets['particleArray'] = np.frombuffer(self.particleArray.tobytes(), dtype=np.uint8)
            np.savez_compressed(location, **{k: v for k, v in datasets.items() if getattr(v, 'dtype', None) != object})
        if printLocation:
            println('Exported arrays to: "{0}".'.format(location))


def _have_h5py():
    try:
        import h5py
        return hasattr(h5py, 'File')          # a namespace stub without the API counts as absent
    except ImportError:
        return False
