"""1-D slab decomposition of the WCSPH step across the GPUs of one box.

One process per GPU (torchrun), one `capi.Context` per process.  Rank r owns the particles with
cut[r] <= x < cut[r+1]; cuts are particle-count quantiles of the fluid so the work is balanced even though
the fluid occupies a corner of the tank.  Per step the ranks exchange, over NCCL (NVLink 5 / NVSwitch):

    all_reduce(MIN)  3 doubles   {h_min, -c_max, -a2_max}            -> identical dt on every rank
    all_gather       12 doubles  counts + local grid bounds          -> message sizes, identical reference grid
    send/recv        migrants (21 doubles each) and halo particles (8 doubles each) with both x-neighbours;
                     halos land directly in the receiver's ghost buffer

The device work between the collectives is the C ABI's slab entry points (csrc/slab.cu); this module only
sequences them.  `comm` is pluggable so the sequencing is unit-tested on CPU with gloo (tests/test_slabs_cpu.py).
"""
import math
import sys

import numpy as np
import torch
import torch.distributed as dist

WIRE_HALO = 8
WIRE_FULL = 21
META = 12


def fluid_quantile_cuts(x_fluid, world):
    """world+1 slab boundaries; the outer ones are infinite so nothing ever leaves the decomposition."""
    cuts = [-math.inf]
    xs = np.sort(np.asarray(x_fluid, dtype=np.float64))
    for r in range(1, world):
        k = min(len(xs) - 1, max(0, (len(xs) * r) // world))
        lo = xs[k - 1] if k > 0 else xs[k]
        cuts.append(float(0.5 * (lo + xs[k])))          # between two particles, never on one
    cuts.append(math.inf)
    for r in range(1, world + 1):                      # degenerate inputs: keep the cuts non-decreasing
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts


def halo_width(kernel, hmax, r0, margin=1.1):
    """Radius inside which a pair can contribute (kernel support or the Lennard-Jones range) plus a margin."""
    q = 3.0 if kernel == 'gaussian' else 2.0
    return max(q * hmax, min(r0, 3.0 * hmax)) * margin


class TorchComm:
    """Collectives through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_min(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)

    def all_gather(self, out, t):
        dist.all_gather_into_tensor(out, t, group=self.group)

    def exchange(self, sends, recvs):
        """sends / recvs: lists of (tensor, peer)."""
        ops = [dist.P2POp(dist.irecv, t, p, self.group) for t, p in recvs] + \
              [dist.P2POp(dist.isend, t, p, self.group) for t, p in sends]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class SlabRun:
    """Sequencer of the slab-decomposed step for one rank."""

    def __init__(self, ctx, comm, cuts, local_pA, local_ids, kernel, r0, hmax, device,
                 mig_frac=0.02, ghost_frac=0.25, min_cap=4096):
        self.ctx, self.comm, self.kernel, self.r0 = ctx, comm, kernel, r0
        self.rank, self.world = comm.rank, comm.world
        self.x_lo, self.x_hi = cuts[self.rank], cuts[self.rank + 1]
        self.left = self.rank - 1 if self.rank > 0 else None
        self.right = self.rank + 1 if self.rank < self.world - 1 else None
        n = len(local_pA)
        self.mig_cap = max(min_cap, int(n * mig_frac))
        self.halo_cap = max(min_cap, int(n * ghost_frac))
        self.ghost_cap = 2 * self.halo_cap + 2 * self.mig_cap
        f64 = dict(dtype=torch.float64, device=device)
        self.ghost = torch.zeros(self.ghost_cap * WIRE_HALO, **f64)
        self.mig_l = torch.zeros(self.mig_cap * WIRE_FULL, **f64)
        self.mig_r = torch.zeros(self.mig_cap * WIRE_FULL, **f64)
        self.mig_in = torch.zeros(2 * self.mig_cap * WIRE_FULL, **f64)
        self.halo_l = torch.zeros(self.halo_cap * WIRE_HALO, **f64)
        self.halo_r = torch.zeros(self.halo_cap * WIRE_HALO, **f64)
        self.meta = torch.zeros(META, **f64)
        self.all_meta = torch.zeros(self.world * META, **f64)
        self.dt3 = torch.zeros(3, **f64)
        self.hmax = float(hmax)
        self.steps = 0
        self.last_counts = None
        ctx.reserve(int(n * 1.3) + self.ghost_cap + 2 * self.mig_cap)
        ctx.upload(local_pA)
        ctx.set_row_ids(local_ids)
        ctx.slab_configure(self.x_lo, self.x_hi, self.ghost.data_ptr(), self.ghost_cap)

    def step(self, nsteps=1, fixed_dt=None, damping=0.0):
        for k in range(nsteps):
            self.ctx.slab_step_plan(k, nsteps)       # corrector of step k fused into the predictor of step k+1
            self._step(fixed_dt, damping)

    def _step(self, fixed_dt, damping):
        ctx, comm = self.ctx, self.comm
        # ---- identical dt on every rank ----
        ctx.slab_dt_local(self.dt3.data_ptr())
        comm.all_reduce_min(self.dt3)
        ctx.slab_step_begin(self.dt3.data_ptr(), fixed_dt, damping)
        # ---- classify + pack, exchange counts and local grid bounds ----
        width = halo_width(self.kernel, self.hmax, self.r0)
        ctx.slab_pack(width, self.mig_l.data_ptr(), self.mig_r.data_ptr(), self.mig_cap,
                      self.halo_l.data_ptr(), self.halo_r.data_ptr(), self.halo_cap, self.meta.data_ptr())
        comm.all_gather(self.all_meta, self.meta)
        M = self.all_meta.cpu().numpy().reshape(self.world, META)       # the one host sync of the step
        if M[:, 10].any():
            raise RuntimeError("slab exchange buffers overflowed (rank(s) %s): raise mig_frac / ghost_frac"
                               % np.flatnonzero(M[:, 10]).tolist())
        me = M[self.rank]
        out_l, out_r, halo_l, halo_r = (int(me[0]), int(me[1]), int(me[2]), int(me[3]))
        if self.left is None:
            assert out_l == 0
        if self.right is None:
            assert out_r == 0
        in_mig_l = int(M[self.left][1]) if self.left is not None else 0
        in_halo_l = int(M[self.left][3]) if self.left is not None else 0
        in_mig_r = int(M[self.right][0]) if self.right is not None else 0
        in_halo_r = int(M[self.right][2]) if self.right is not None else 0
        bounds = M[:, 4:10].min(axis=0)
        self.hmax = float(-bounds[5])
        n_own_ghost = out_l + out_r
        n_ghost = n_own_ghost + in_halo_l + in_halo_r
        if n_ghost > self.ghost_cap or in_mig_l + in_mig_r > 2 * self.mig_cap:
            raise RuntimeError("slab receive buffers too small")
        # ---- payloads ----
        sends, recvs = [], []
        g0 = n_own_ghost * WIRE_HALO
        if self.left is not None:
            if out_l: sends.append((self.mig_l[:out_l * WIRE_FULL], self.left))
            if halo_l: sends.append((self.halo_l[:halo_l * WIRE_HALO], self.left))
            if in_mig_l: recvs.append((self.mig_in[:in_mig_l * WIRE_FULL], self.left))
            if in_halo_l: recvs.append((self.ghost[g0:g0 + in_halo_l * WIRE_HALO], self.left))
        g1 = g0 + in_halo_l * WIRE_HALO
        m1 = in_mig_l * WIRE_FULL
        if self.right is not None:
            if out_r: sends.append((self.mig_r[:out_r * WIRE_FULL], self.right))
            if halo_r: sends.append((self.halo_r[:halo_r * WIRE_HALO], self.right))
            if in_mig_r: recvs.append((self.mig_in[m1:m1 + in_mig_r * WIRE_FULL], self.right))
            if in_halo_r: recvs.append((self.ghost[g1:g1 + in_halo_r * WIRE_HALO], self.right))
        comm.exchange(sends, recvs)
        # ---- owned set update, same grid everywhere, then the force evaluation and the corrector ----
        ctx.slab_commit(out_l + out_r, self.mig_in.data_ptr(), in_mig_l + in_mig_r, n_ghost, bounds.tolist())
        ctx.slab_step_end(damping)
        self.steps += 1
        self.last_counts = dict(mig_out=(out_l, out_r), halo_out=(halo_l, halo_r), mig_in=(in_mig_l, in_mig_r),
                                halo_in=(in_halo_l, in_halo_r), ghosts=n_ghost, owned=ctx.num_active)

    def export_device(self, fields):
        """(ids, labels, [columns]) of the owned particles as tensors on the run's device."""
        return _export_device(self.ctx, self.ghost.device, fields)

    def export(self, fields):
        """(ids, labels, {field: column}) of the owned particles, as host numpy arrays."""
        ids, lab, cols = self.export_device(fields)
        return ids.cpu().numpy(), lab.cpu().numpy(), {f: c.cpu().numpy() for f, c in zip(fields, cols)}

    def set_bounds(self, x_lo, x_hi):
        self.x_lo, self.x_hi = float(x_lo), float(x_hi)
        self.ctx.slab_configure(self.x_lo, self.x_hi, self.ghost.data_ptr(), self.ghost_cap)


class NcclSlabRun:
    """Same protocol as SlabRun, sequenced inside libosph_b200 with direct NCCL calls (osph_slab_run).
    torch.distributed only distributes the 128-byte NCCL unique id."""

    def __init__(self, ctx, cuts, local_pA, local_ids, kernel, r0, hmax, device, group=None,
                 mig_frac=0.02, ghost_frac=0.25, min_cap=4096):
        from osph_b200 import capi
        self.ctx, self.kernel, self.r0 = ctx, kernel, r0
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.x_lo, self.x_hi = cuts[self.rank], cuts[self.rank + 1]
        self.device = device
        n = len(local_pA)
        self.mig_cap = max(min_cap, int(n * mig_frac))
        self.halo_cap = max(min_cap, int(n * ghost_frac))
        self.ghost_cap = 2 * self.halo_cap + 2 * self.mig_cap
        uid = torch.zeros(128, dtype=torch.uint8, device=device)
        if self.rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0, group=group)
        ctx.reserve(int(n * 1.3) + self.ghost_cap + 2 * self.mig_cap)
        ctx.upload(local_pA)
        ctx.set_row_ids(local_ids)
        self.comm = ctx.slab_comm_create(bytes(uid.cpu().numpy().tobytes()), self.rank, self.world, self.x_lo, self.x_hi,
                                         r0, float(hmax), self.mig_cap, self.halo_cap)
        self.steps = 0

    def step(self, nsteps=1, fixed_dt=None, damping=0.0):
        self.ctx.slab_run(self.comm, nsteps, fixed_dt, damping)
        self.steps += nsteps

    @property
    def last_counts(self):
        c = self.ctx.slab_last_counts(self.comm)
        return dict(mig_out=(c[0], c[1]), halo_out=(c[2], c[3]), mig_in=(c[4], c[5]), halo_in=(c[6], c[7]),
                    ghosts=c[0] + c[1] + c[6] + c[7], owned=self.ctx.num_active)

    def reattach(self):
        self.ctx.slab_comm_attach(self.comm)

    def export_device(self, fields):
        return _export_device(self.ctx, self.device, fields)

    def export(self, fields):
        ids, lab, cols = self.export_device(fields)
        return ids.cpu().numpy(), lab.cpu().numpy(), {f: c.cpu().numpy() for f, c in zip(fields, cols)}

    def set_bounds(self, x_lo, x_hi):
        self.x_lo, self.x_hi = float(x_lo), float(x_hi)
        self.ctx.slab_comm_set_bounds(self.comm, self.x_lo, self.x_hi)

    def close(self):
        if self.comm is not None:
            self.ctx.slab_comm_destroy(self.comm)
            self.comm = None


class P2PSlabRun(NcclSlabRun):
    """Slab protocol over NVLink peer memory (osph_slab_p2p_run): IPC windows, the pack kernel writes into the
    neighbours' HBM, mailbox kernels instead of NCCL collectives.  torch.distributed only all-gathers the 64-byte
    IPC handles at start-up."""

    def __init__(self, ctx, cuts, local_pA, local_ids, kernel, r0, hmax, device, group=None,
                 mig_frac=0.02, ghost_frac=0.25, min_cap=4096):
        self.ctx, self.kernel, self.r0 = ctx, kernel, r0
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.x_lo, self.x_hi = cuts[self.rank], cuts[self.rank + 1]
        self.device = device
        n = len(local_pA)
        self.mig_cap = max(min_cap, int(n * mig_frac))
        self.halo_cap = max(min_cap, int(n * ghost_frac))
        # every rank addresses its neighbours' windows with its OWN layout, so the region capacities must be identical
        caps = torch.tensor([self.mig_cap, self.halo_cap], dtype=torch.int64, device=device)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)
        self.mig_cap, self.halo_cap = int(caps[0]), int(caps[1])
        self.ghost_cap = 2 * self.halo_cap + 2 * self.mig_cap
        ctx.reserve(int(n * 1.3) + self.ghost_cap + 2 * self.mig_cap)
        ctx.upload(local_pA)
        ctx.set_row_ids(local_ids)
        self.comm, handle = ctx.slab_p2p_create(self.rank, self.world, self.x_lo, self.x_hi, r0, float(hmax),
                                                self.mig_cap, self.halo_cap)
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(device)
        allh = torch.zeros(64 * self.world, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        ctx.slab_p2p_connect(self.comm, allh.cpu().numpy().tobytes())
        dist.barrier(group=group)
        self.steps = 0

    def step(self, nsteps=1, fixed_dt=None, damping=0.0):
        self.ctx.slab_p2p_run(self.comm, nsteps, fixed_dt, damping)
        self.steps += nsteps

    @property
    def last_counts(self):
        c = self.ctx.slab_p2p_last_counts(self.comm)
        return dict(mig_out=(c[0], c[1]), halo_out=(c[2], c[3]), mig_in=(c[4], c[5]), halo_in=(c[6], c[7]),
                    ghosts=c[0] + c[1] + c[6] + c[7], owned=self.ctx.num_active)

    @property
    def cadence_stats(self):
        """(steps that sorted, steps that reused the binning and the frozen halo lists)."""
        return self.ctx.slab_p2p_stats(self.comm)

    def reattach(self):
        self.ctx.slab_p2p_attach(self.comm)

    def set_bounds(self, x_lo, x_hi):
        self.x_lo, self.x_hi = float(x_lo), float(x_hi)
        self.ctx.slab_p2p_set_bounds(self.comm, self.x_lo, self.x_hi)

    def close(self):
        if self.comm is not None:
            dist.barrier()                       # nobody may still be writing into a window that is about to go away
            self.ctx.slab_p2p_destroy(self.comm)
            self.comm = None


def _export_device(ctx, device, fields):
    n = ctx.num_active
    ids = torch.zeros(max(n, 1), dtype=torch.int32, device=device)
    lab = torch.zeros(max(n, 1), dtype=torch.int8, device=device)
    cols = [torch.zeros(max(n, 1), dtype=torch.float64, device=device) for _ in fields]
    ctx.slab_export(ids.data_ptr(), lab.data_ptr(), fields, [c.data_ptr() for c in cols])
    return ids[:n], lab[:n], [c[:n] for c in cols]


def rebalance(run, group=None, bins=4096, max_move_frac=0.5):
    """Re-cut the slabs so every rank owns the same number of fluid particles again (the fluid of a dam break
    leaves the corner it started in; SURVEY.md section 7, hard part 7).

    Collective: every rank must call it at the same step.  A histogram of the fluid x-coordinates over `bins`
    columns is summed over the ranks (one all_reduce of `bins` doubles); the new cuts are its quantiles.  A cut
    moves at most `max_move_frac` of the way into the neighbouring slab, so the particles between the old and the
    new cut reach their new owner through the ordinary migration of the next step (migrants only ever travel to the
    adjacent rank).  Returns the new list of cuts."""
    world, rank = run.world, run.rank
    _, lab, (x,) = run.export_device(['x'])
    fx = x[lab == 0]
    dev = x.device
    big = 1e300
    ext = torch.tensor([fx.min().item() if fx.numel() else big, -(fx.max().item() if fx.numel() else -big)],
                       dtype=torch.float64, device=dev)
    dist.all_reduce(ext, op=dist.ReduceOp.MIN, group=group)
    lo, hi = float(ext[0]), float(-ext[1])
    if not (hi > lo):
        return None
    hist = torch.histc(fx, bins=bins, min=lo, max=hi).to(torch.float64) if fx.numel() else \
        torch.zeros(bins, dtype=torch.float64, device=dev)
    # current inner cuts travel along so that every rank limits the moves identically
    mine = torch.full((world + 1,), big, dtype=torch.float64, device=dev)
    mine[rank] = run.x_lo if math.isfinite(run.x_lo) else -big
    mine[rank + 1] = min(float(mine[rank + 1]), run.x_hi if math.isfinite(run.x_hi) else big)
    payload = torch.cat((hist, mine))
    red = payload.clone()
    dist.all_reduce(red[:bins], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(red[bins:], op=dist.ReduceOp.MIN, group=group)
    hist = red[:bins].cpu().numpy()
    old = red[bins:].cpu().numpy()
    cdf = np.cumsum(hist)
    total = cdf[-1]
    if total <= 0:
        return None
    width = (hi - lo) / bins
    cuts = [-math.inf]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cdf, target))
        b = min(b, bins - 1)
        before = cdf[b - 1] if b > 0 else 0.0
        frac = (target - before) / hist[b] if hist[b] > 0 else 0.0
        new = lo + (b + frac) * width
        # stay inside the two slabs that meet at this cut
        left_lo = old[r - 1] if r - 1 > 0 else lo
        right_hi = old[r + 1] if r + 1 < world else hi
        cur = old[r]
        new = min(max(new, cur - max_move_frac * (cur - left_lo)), cur + max_move_frac * (right_hi - cur))
        cuts.append(float(max(new, cuts[-1])))
    cuts.append(math.inf)
    run.set_bounds(cuts[rank], cuts[rank + 1])
    return cuts


def partition(pA, world, rank):
    """Slab cuts from the fluid quantiles and this rank's rows of the global array."""
    act = ~pA['deleted']
    fluid = act & (pA['label'] == 0)
    cuts = fluid_quantile_cuts(pA['x'][fluid] if fluid.any() else pA['x'][act], world)
    mine = act & (pA['x'] >= cuts[rank]) & (pA['x'] < cuts[rank + 1])
    ids = np.flatnonzero(mine).astype(np.int32)
    return cuts, np.ascontiguousarray(pA[ids]), ids


def gather_global(run, pA_template, fields, group=None):
    """Assemble the global particle array on every rank from the owned sets (validation helper)."""
    ids, lab, cols = run.export(fields)
    parts = [None] * run.world
    dist.all_gather_object(parts, (ids, cols), group=group)
    out = pA_template.copy()
    seen = np.zeros(len(out), dtype=np.int64)
    for pid, pcols in parts:
        seen[pid] += 1
        for f in fields:
            out[f][pid] = pcols[f]
    return out, seen


# ------------------------------------------------------------------------------------------------
# multi-GPU arm of bench.py
# ------------------------------------------------------------------------------------------------
def bench_multi_gpu(args, rank, world, local, cpu_baseline=None):
    """cpu_baseline: callable(case) -> dict, supplied by bench.py (the CPU checker is test infrastructure and is not
    known to this package); called on rank 0 only, outside the timed regions."""
    import json
    import os
    import time
    from osph_b200 import capi
    import bench as B

    prec = capi.FP64 if args.precision == "fp64" else capi.FP32
    F = 8 if prec == capi.FP64 else 4
    n_side, scaling = B.resolve_problem(args, world)
    B.WORKLOAD = getattr(args, 'workload', 'dam_break')
    case = B.build_case(n_side)
    pA, c = case['pA'], case['consts']
    n_total = len(pA)
    hmax0 = case['h'] if case['h'] is not None else float(pA['h'].max())
    cuts, local_pA, ids = partition(pA, world, rank)
    del pA
    cfg = capi.make_config(c, args.kernel, 'pec', prec, case['h'], device=local)
    ctx = capi.Context(cfg)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    torch.cuda.set_stream(stream)
    if args.sequencer == 'python':
        run = SlabRun(ctx, TorchComm(), cuts, local_pA, ids, args.kernel, case['r0'], hmax0, torch.device('cuda', local))
    elif args.sequencer == 'nccl':
        run = NcclSlabRun(ctx, cuts, local_pA, ids, args.kernel, case['r0'], hmax0, torch.device('cuda', local))
    else:
        # default: NVLink peer-memory exchange; if CUDA IPC is not usable between the ranks (e.g. GPUs without peer
        # access), every rank falls back to the NCCL sequencer together
        run, err = None, None
        try:
            run = P2PSlabRun(ctx, cuts, local_pA, ids, args.kernel, case['r0'], hmax0, torch.device('cuda', local))
        except Exception as e:      # noqa: BLE001
            err = e
        ok = torch.tensor([1 if run is not None else 0], device='cuda')
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            if run is not None:
                run.close()
            if rank == 0:
                print("bench: peer-memory sequencer unavailable (%s); using NCCL" % err, file=sys.stderr)
            args.sequencer = 'nccl'
            run = NcclSlabRun(ctx, cuts, local_pA, ids, args.kernel, case['r0'], hmax0, torch.device('cuda', local))

    clocks = B.ClockSampler(local) if rank == 0 else None
    run.step(args.warmup, None, B.DAMPING)
    ctx.sync(); ctx.pair_kernel_time()
    l0 = ctx.launch_count
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if clocks:
        clocks.mark_start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run.step(args.steps, None, B.DAMPING)
    e1.record(stream)
    e1.synchronize(); torch.cuda.synchronize()
    if clocks:
        clocks.mark_end()
    t_local = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device='cuda')
    dist.barrier()
    dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    t_dev = float(t_local.item())
    clk = clocks.stop() if clocks else None
    launches = ctx.launch_count - l0
    pair_us, pair_n = ctx.pair_kernel_time()
    owned = torch.tensor([ctx.num_active, run.last_counts['ghosts'], launches], dtype=torch.float64, device='cuda')
    allowned = torch.zeros(world * 3, dtype=torch.float64, device='cuda')
    dist.all_gather_into_tensor(allowned, owned)
    pair_t = torch.tensor([pair_us], dtype=torch.float64, device='cuda')
    dist.all_reduce(pair_t, op=dist.ReduceOp.MAX)

    # ---- end to end: host buffers every step (upload slab, one step, download slab) ----
    e2e_steps = max(3, min(args.steps, 10))
    n_loc = ctx.num_active
    cap_rows = int(n_loc * 1.1) + 1024
    hbuf = torch.empty(cap_rows * 154, dtype=torch.uint8, pin_memory=True)
    host = hbuf.numpy().view(local_pA.dtype)
    hids = torch.empty(cap_rows, dtype=torch.int32, pin_memory=True).numpy()
    n_loc = ctx.download_owned(host, hids)

    def e2e_step(n_rows):
        ctx.upload(host[:n_rows])
        ctx.set_row_ids(hids[:n_rows])
        if args.sequencer == 'python':
            ctx.slab_configure(run.x_lo, run.x_hi, run.ghost.data_ptr(), run.ghost_cap)
        else:
            run.reattach()
        run.step(1, None, B.DAMPING)
        return ctx.download_owned(host, hids)

    n_loc = e2e_step(n_loc)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    moved = 0
    for _ in range(e2e_steps):
        moved += n_loc * 154
        n_loc = e2e_step(n_loc)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    mv = torch.tensor([moved / e2e_steps], dtype=torch.float64, device='cuda')
    dist.all_reduce(mv, op=dist.ReduceOp.SUM)

    if rank == 0:
        peaks, which = B.measured_peaks()
        per = allowned.cpu().numpy().reshape(world, 3)
        # DRAM bytes of one launch: the single-GPU ncu capture of the same kernel, per particle it walks, times the
        # particles (owned + ghosts) the slowest rank's launch walks
        traffic = None
        try:
            with open(os.path.join(B.ROOT, "profiles", "pair_kernel_traffic.json")) as f:
                t = json.load(f).get("%s_%s_N1000" % (args.precision, args.kernel))
            if t:
                traffic = int(t["dram_bytes_per_launch"] / t.get("particles", 1009603) * float(per[:, 0].max() + per[rank, 1]))
        except Exception:       # noqa: BLE001
            pass
        cpu = cpu_baseline(case) if cpu_baseline is not None else None
        n_pair = float(per[:, 0].max() + per[rank, 1])
        alg_bytes = (13 * F + 1) * float(per[:, 0].max())
        pu = float(pair_t.item())
        achieved = alg_bytes / (pu * 1e-6) / 1e9 if pu > 0 else 0.0
        line = {
            "metric": B.METRIC, "value": n_total * args.steps / t_dev, "unit": B.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64" if prec == capi.FP64 else "f32 pair arithmetic, f64 state", "data": "synthetic",
            "config": {"workload": B.workload_name(n_side, n_total, args.kernel, args.precision.upper()),
                       "particles": n_total, "particles_per_gpu": [int(v) for v in per[:, 0]],
                       "ghosts_per_gpu": [int(v) for v in per[:, 1]], "damping": B.DAMPING, "dt": "dynamic",
                       "l2": "per-GPU working set exceeds the 126 MB L2",
                       "parallelism": "1-D slabs along x, %d ranks, halo + migration over %s, fluid-quantile cuts" % (
                           world, {"p2p": "NVLink peer-memory windows (mailbox kernels, pack kernel writes into peers)",
                                   "nccl": "NCCL send/recv sequenced in C++", "python": "NCCL via torch.distributed"}[args.sequencer])},
            "clocks": clk, "gpu_launches": int(per[:, 2].sum()),
            "e2e": {"value": n_total * e2e_steps / float(t_e2e.item()), "unit": B.UNIT,
                    "h2d_bytes_per_step": int(mv.item()), "d2h_bytes_per_step": int(mv.item()), "steps": e2e_steps,
                    "what": "per rank: osph_upload_aos(pinned slab) + one slab step + download of the owned records"},
            "roofline": {"bound": "hbm", "kernel": "k_pair (slowest rank)", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "traffic_source": "single-GPU ncu capture scaled by the particles this launch walks",
                         "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)", "avg_launch_us": pu,
                         "share_of_step": pu * 1e-6 * args.steps / t_dev,
                         "algorithmic_bytes_per_particle": 13 * F + 1, "pair_kernel_particles": n_pair},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    dist.barrier()
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())       # never leave torch on a stream about to be destroyed
    if hasattr(run, 'close'):
        run.close()
    ctx.close()
    dist.destroy_process_group()
