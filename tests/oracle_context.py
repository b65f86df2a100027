"""A CPU stand-in for osph_b200.capi.Context built on the oracle -- TEST INFRASTRUCTURE ONLY.

It implements the handful of Context methods the re-hosted Solver (offshore-sph_b200/src/Solver.py) calls, with the
CPU restatement of the reference doing the arithmetic, so that the Solver's HOST logic -- settling, gate removal, masks,
export pipeline, coupling round trips, timing -- is exercised by `pytest -m "not gpu"` against the reference Solver's
golden end state.  The product never imports this file: tests monkeypatch `capi.Context` with it.
"""
import numpy as np

from oracle import oracle as O

KERNEL_NAMES = {0: 'cubic', 1: 'wendland', 2: 'gaussian'}
STATE = ('m', 'rho', 'p', 'c', 'drho', 'h', 'x', 'y', 'vx', 'vy', 'ax', 'ay', 'xsphx', 'xsphy', 'x0', 'y0', 'vx0', 'vy0',
         'rho0')


class OracleContext:
    instances = []

    def __init__(self, cfg):
        self.cfg = cfg
        self.w = O.wcsph(cfg.height, cfg.r0, cfg.rho0, bool(cfg.method_xsph), cfg.Pb, bool(cfg.summation_density))
        assert self.w.co == cfg.co and self.w.B == cfg.B
        self.kernel = KERNEL_NAMES[int(cfg.kernel)]
        self.P = None
        self.grid = None
        self.calls = {}
        self._tickets = {}
        self._next_ticket = 0
        OracleContext.instances.append(self)

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # ---- transfers -----------------------------------------------------------------------------------------------
    def upload(self, pA):
        self._count('upload')
        assert pA.flags['C_CONTIGUOUS'] and pA.dtype.itemsize == 154
        self.mirror = pA.copy()
        self.act = ~pA['deleted']
        self.P = O.Particles.from_aos(pA[self.act])
        self.fluid = self.P.fluid.astype(np.uint8)
        self.grid = None

    def download(self, pA):
        self._count('download')
        for f in STATE:
            pA[f][self.act] = getattr(self.P, f)
        return pA

    def download_fields(self, names):
        self._count('download_fields')
        return {f: getattr(self.P, f).copy() for f in names}

    def export_begin(self, names, rows=False):
        self._count('export_begin')
        assert len(self._tickets) < 2, "both export slots in flight"
        cols = {}
        for f in names:
            if rows:
                col = self.mirror[f].copy()
                col[self.act] = getattr(self.P, f)
            else:
                col = getattr(self.P, f).copy()
            cols[f] = col
        t = self._next_ticket
        self._next_ticket += 1
        self._tickets[t] = cols
        return (t, tuple(names), len(self.mirror) if rows else self.P.n)

    def export_end(self, ticket):
        self._count('export_end')
        return self._tickets.pop(ticket[0])

    def _slots(self, rows):
        rank = np.cumsum(self.act) - 1
        rows = np.asarray(rows, dtype=np.int64)
        assert np.all(self.act[rows]), "row transfer of a deleted row"
        return rank[rows]

    def download_rows(self, rows, pA):
        self._count('download_rows')
        s = self._slots(rows)
        for f in STATE:
            pA[f][rows] = getattr(self.P, f)[s]
        return pA

    def upload_rows(self, rows, pA):
        self._count('upload_rows')
        s = self._slots(rows)
        for f in STATE:
            getattr(self.P, f)[s] = pA[f][rows]
        self.grid = None

    def set_active(self, active):
        """Rows switched off / on: state back into the record mirror, new deleted flags, active set rebuilt."""
        self._count('set_active')
        active = np.asarray(active) != 0
        assert len(active) == len(self.mirror)
        for f in STATE:
            self.mirror[f][self.act] = getattr(self.P, f)
        self.mirror['deleted'] = ~active
        self.act = active.copy()
        self.P = O.Particles.from_aos(self.mirror[self.act])
        self.fluid = self.P.fluid.astype(np.uint8)
        self.grid = None

    @property
    def num_active(self):
        return self.P.n

    @property
    def num_fluid(self):
        return int(self.fluid.sum())

    # ---- the step ------------------------------------------------------------------------------------------------
    def _refresh_h(self):
        P, fl = self.P, self.fluid.astype(bool)
        if int(self.cfg.dynamic_h) == 0:
            P.h[fl] = self.cfg.fixed_h
        elif int(self.cfg.dynamic_h) == 1:
            P.h[fl] = O.compute_h(self.cfg.h_sigma, P.m[fl], P.rho[fl])

    def initialize(self):
        """reference src/Solver.py:184-196"""
        P, fl = self.P, self.fluid.astype(bool)
        self._refresh_h()
        P.rho[fl] = O.initialize_density(self.w, P.y[fl])
        P.p[fl] = ((P.rho[fl] / self.cfg.rho0) ** self.cfg.gamma - 1.0) * self.cfg.B + self.cfg.Pb
        P.c[fl] = self.cfg.co

    def timestep(self):
        return O.timestep(self.P, self.fluid)

    def predict(self, dt, damping):
        integ = int(self.cfg.integrator)
        if integ == 0:
            O.pec_predict(self.P, self.fluid, dt, damping, bool(self.cfg.integrator_xsph), bool(self.cfg.strict))
        elif integ == 2:
            O.verlet_predict(self.P, self.fluid, dt)
        self.grid = None

    def build_neighbours(self):
        self.grid = O.Grid(self.P, self.cfg.nn_scale)            # the grid sees h BEFORE the refresh (Solver.py:238-246)
        self._refresh_h()

    def compute(self):
        if self.grid is None:
            self.build_neighbours()
        O.loop(self.P, self.w, self.grid, self.kernel)

    def correct(self, dt, damping):
        integ = int(self.cfg.integrator)
        if integ == 0:
            O.pec_correct(self.P, self.fluid, dt, damping, bool(self.cfg.integrator_xsph), bool(self.cfg.strict))
        elif integ == 1:
            O.euler_correct(self.P, self.fluid, dt)
        else:
            O.verlet_correct(self.P, self.fluid, dt, bool(self.cfg.integrator_xsph))
        self.grid = None

    def kinetic_energy(self):
        return O.kinetic_energy(self.P, self.fluid)

    def probe_pressure(self, x, y, h):
        raise NotImplementedError("device pressure probe: GPU tests only")

    def sync(self):
        return 0

    def timers(self):
        return {}

    def close(self):
        self.P = None
