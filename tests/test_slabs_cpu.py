"""CPU (gloo, world_size 2 and 3): the slab sequencer of osph_b200/slabs.py -- quantile cuts, message sizes,
buffer offsets, migration, ghost completeness -- against a brute-force global computation.

The device side is replaced by a numpy stand-in with the same entry points as the C ABI's slab calls
(tests only; the product path has no CPU implementation).  Its "physics": particles drift with constant
velocity, and the force evaluation counts, for every owned particle, the owned + ghost particles within the
interaction radius.  If a halo or a migrant were lost, misplaced or duplicated, the counts would differ from the
global brute force.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from osph_b200 import slabs
from oracle.oracle import particle_dtype

R = 0.11          # interaction radius = halo width passed by the sequencer (kernel 'cubic': 2 * hmax * 1.1)
HMAX = 0.05
DT = 0.05


def _view(ptr, n):
    return np.ctypeslib.as_array((C.c_double * n).from_address(ptr)) if n > 0 else np.zeros(0)


class MockContext:
    """numpy stand-in for capi.Context's slab entry points (test infrastructure)."""

    def reserve(self, cap): self.cap = cap

    def slab_step_plan(self, step, nsteps): pass

    def upload(self, pA):
        self.x = pA['x'].copy(); self.y = pA['y'].copy(); self.vx = pA['vx'].copy(); self.count = np.zeros(len(pA))

    def set_row_ids(self, ids): self.ids = np.asarray(ids, dtype=np.int64).copy()

    def slab_configure(self, lo, hi, ghost_ptr, ghost_cap):
        self.lo, self.hi, self.ghost_ptr, self.ghost_cap = lo, hi, ghost_ptr, ghost_cap

    @property
    def num_active(self): return len(self.x)

    def slab_dt_local(self, ptr): _view(ptr, 3)[:] = [DT * (1 + 0.01 * len(self.x)), 0.0, 0.0]   # rank-dependent on purpose

    def slab_step_begin(self, ptr, fixed_dt, damping):
        self.dt = _view(ptr, 3)[0]
        self.x = self.x + self.vx * self.dt

    def slab_pack(self, width, mig_l, mig_r, mig_cap, halo_l, halo_r, halo_cap, meta_ptr):
        assert abs(width - R) < 1e-12
        ghost = _view(self.ghost_ptr, self.ghost_cap * 8).reshape(-1, 8)
        left = self.x < self.lo; right = self.x >= self.hi
        def full(sel):
            rec = np.zeros((sel.sum(), 21)); rec[:, 6] = self.x[sel]; rec[:, 7] = self.y[sel]; rec[:, 8] = self.vx[sel]
            rec[:, 20] = self.ids[sel]; return rec
        def light(sel):
            rec = np.zeros((sel.sum(), 8)); rec[:, 0] = self.x[sel]; rec[:, 1] = self.y[sel]; rec[:, 2] = self.vx[sel]; return rec
        ml, mr = full(left), full(right)
        _view(mig_l, mig_cap * 21).reshape(-1, 21)[:len(ml)] = ml
        _view(mig_r, mig_cap * 21).reshape(-1, 21)[:len(mr)] = mr
        own = np.concatenate((light(left), light(right)))
        ghost[:len(own)] = own
        stay = ~(left | right)
        hl = light(stay & (self.x < self.lo + width)); hr = light(stay & (self.x >= self.hi - width))
        _view(halo_l, halo_cap * 8).reshape(-1, 8)[:len(hl)] = hl
        _view(halo_r, halo_cap * 8).reshape(-1, 8)[:len(hr)] = hr
        self._stay = stay
        big = 1e300
        _view(meta_ptr, 12)[:] = [len(ml), len(mr), len(hl), len(hr),
                                  self.x.min() if len(self.x) else big, -(self.x.max() if len(self.x) else -big),
                                  0.0, -1.0, HMAX, -HMAX, 0, 0]

    def slab_commit(self, n_out, mig_in_ptr, n_in, n_ghost, bounds):
        assert n_out == int((~self._stay).sum())
        rec = _view(mig_in_ptr, n_in * 21).reshape(-1, 21)
        self.x = np.concatenate((self.x[self._stay], rec[:, 6])); self.y = np.concatenate((self.y[self._stay], rec[:, 7]))
        self.vx = np.concatenate((self.vx[self._stay], rec[:, 8]))
        self.ids = np.concatenate((self.ids[self._stay], rec[:, 20].astype(np.int64)))
        self.n_ghost = n_ghost; self.bounds = np.asarray(bounds)

    def slab_step_end(self, damping):
        g = _view(self.ghost_ptr, self.n_ghost * 8).reshape(-1, 8)
        ax = np.concatenate((self.x, g[:, 0])); ay = np.concatenate((self.y, g[:, 1]))
        d2 = (self.x[:, None] - ax[None, :]) ** 2 + (self.y[:, None] - ay[None, :]) ** 2
        self.count = (d2 <= (R / 1.1) ** 2).sum(axis=1).astype(np.float64)

    def slab_export(self, ids_ptr, label_ptr, fields, col_ptrs):
        n = len(self.x)
        np.ctypeslib.as_array((C.c_int32 * n).from_address(ids_ptr))[:] = self.ids
        for f, p in zip(fields, col_ptrs):
            _view(p, n)[:] = {'x': self.x, 'ax': self.count}[f]


def _particles(n=1500, seed=0):
    rng = np.random.default_rng(seed)
    pA = np.zeros(n, dtype=particle_dtype)
    pA['x'] = rng.uniform(0, 3, n); pA['y'] = rng.uniform(0, 0.5, n)
    pA['vx'] = rng.uniform(-1, 1, n)          # up to 0.05 per step: crossings every step, < halo width
    pA['label'][n - 50:] = 1                    # some non-fluid rows: cuts use the fluid only
    return pA


def _worker_rebalance(rank, world, port, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pA = _particles()
        pA['label'] = 0
        pA['vx'] = 0.04 * pA['x']                       # the cloud stretches to the right: static cuts would starve rank 0
        cuts, local, ids = slabs.partition(pA, world, rank)
        run = slabs.SlabRun(MockContext(), slabs.TorchComm(), cuts, local, ids, 'cubic', 0.01, HMAX,
                            torch.device('cpu'), min_cap=4096)
        hist = []
        for k in range(steps):
            run.step(1, None, 0.0)
            if k % 2 == 1 and k < steps - 1:              # the step after a re-cut carries out its migration
                slabs.rebalance(run, bins=512)
            hist.append(run.ctx.num_active)
        out, seen = slabs.gather_global(run, pA, ['x', 'ax'])
        q.put((rank, hist, out['x'].copy(), out['ax'].copy(), seen, np.asarray([run.x_lo, run.x_hi])))
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pA = _particles()
        cuts, local, ids = slabs.partition(pA, world, rank)
        run = slabs.SlabRun(MockContext(), slabs.TorchComm(), cuts, local, ids, 'cubic', 0.01, HMAX,
                            torch.device('cpu'), min_cap=2048)
        dts = []
        for _ in range(steps):
            run.step(1, None, 0.0)
            dts.append(run.ctx.dt)
        out, seen = slabs.gather_global(run, pA, ['x', 'ax'])
        q.put((rank, cuts, np.asarray(dts), out['x'].copy(), out['ax'].copy(), seen, run.last_counts,
               run.ctx.bounds.copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world", [2, 3])
def test_slab_sequencer_matches_global_brute_force(world):
    steps = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, q)) for r in range(world)]
    [p.start() for p in procs]
    res = []
    import queue, time
    t_end = time.time() + 180
    while len(res) < world and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == world and all(p.exitcode == 0 for p in procs)
    res.sort(key=lambda t: t[0])

    pA = _particles()
    rank0 = res[0]
    cuts, dts = rank0[1], rank0[2]
    assert len(cuts) == world + 1 and cuts[0] == -np.inf and cuts[-1] == np.inf and cuts == sorted(cuts)
    fl = pA['x'][pA['label'] == 0]
    per = [np.sum((fl >= cuts[r]) & (fl < cuts[r + 1])) for r in range(world)]
    assert max(per) - min(per) <= 2                                    # fluid-balanced cuts
    for r in res:                                                      # identical dt on every rank (MIN all-reduce)
        assert np.array_equal(r[2], dts)
    x = pA['x'].copy()
    for dt in dts:
        x = x + pA['vx'] * dt
    d2 = (x[:, None] - x[None, :]) ** 2 + (pA['y'][:, None] - pA['y'][None, :]) ** 2
    want = (d2 <= (R / 1.1) ** 2).sum(axis=1).astype(np.float64)
    for r in res:
        assert np.array_equal(r[5], np.ones(len(pA), dtype=np.int64))  # every particle owned exactly once
        assert np.allclose(r[3], x, rtol=0, atol=1e-15)                # migrants kept their state and identity
        assert np.array_equal(r[4], want)                              # no neighbour lost at the slab faces
        assert r[7][0] == x.min() and -r[7][1] == x.max()              # same global grid bounds everywhere
    assert sum(sum(r[6]['mig_out']) for r in res) == sum(sum(r[6]['mig_in']) for r in res)
    assert sum(sum(r[6]['mig_out']) for r in res) > 0                  # the test did exercise migration


def test_quantile_cuts_and_halo_width():
    cuts = slabs.fluid_quantile_cuts(np.arange(100.0), 4)
    assert cuts[1:-1] == [24.5, 49.5, 74.5]
    assert slabs.fluid_quantile_cuts(np.zeros(10), 3)[1:-1] == [0.0, 0.0]        # degenerate: non-decreasing
    assert slabs.halo_width('cubic', 0.5, 0.25) == pytest.approx(1.1)
    assert slabs.halo_width('gaussian', 0.5, 0.25) == pytest.approx(1.65)
    assert slabs.halo_width('cubic', 0.1, 1.0) == pytest.approx(0.33)          # Lennard-Jones range min(r0, 3h) wins


def _spawn(target, world, steps):
    import queue, time
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, steps, q)) for r in range(world)]
    [p.start() for p in procs]
    res, t_end = [], time.time() + 180
    while len(res) < world and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == world and all(p.exitcode == 0 for p in procs)
    return sorted(res, key=lambda t: t[0])


def test_rebalance_keeps_the_load_even_and_the_physics_intact():
    world, steps = 3, 13
    res = _spawn(_worker_rebalance, world, steps)
    pA = _particles(); pA['label'] = 0; pA['vx'] = 0.04 * pA['x']
    n = len(pA)
    final_counts = [r[1][-1] for r in res]
    assert sum(final_counts) == n
    # the cloud stretches by ~15 % over the run: static cuts would end near [435, 435, 630]; re-cut slabs stay even
    assert max(final_counts) - min(final_counts) < 0.06 * n / world
    for r in res:
        assert np.array_equal(r[4], np.ones(n, dtype=np.int64))            # every particle owned exactly once
    # slabs tile the line and the physics (neighbour counts through ghosts) is still the global brute force
    edges = sorted(set(np.concatenate([r[5] for r in res]).tolist()))
    assert edges[0] == -np.inf and edges[-1] == np.inf and len(edges) == world + 1
    x = res[0][2]
    d2 = (x[:, None] - x[None, :]) ** 2 + (pA['y'][:, None] - pA['y'][None, :]) ** 2
    want = (d2 <= (R / 1.1) ** 2).sum(axis=1).astype(np.float64)
    for r in res:
        assert np.array_equal(r[3], want)
