"""CPU (gloo, world_size 2 and 3): the slab-decomposed step with the REAL device code of csrc/slab.cu, step.cu and
pair.cu executed under the SIMT emulator (tests/emu), sequenced by osph_b200.slabs.SlabRun over torch.distributed.

Same assertions as tests/multi_gpu_check.py makes on the GPUs: after a run with migration across the slab faces and
a mid-run re-cut of the slabs, the gathered result equals the single-rank run to summation order, every particle is
owned exactly once, and dt is identical on every rank.  Test infrastructure only (see tests/emu/emu.h).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build as emu_build  # noqa: E402

if not emu_build.available():
    pytest.skip("g++ or the CUDA headers are missing: cannot build the emulated library", allow_module_level=True)

FIXED_DT = 2e-4
FIELDS = ['x', 'y', 'vx', 'vy', 'rho', 'p', 'drho', 'ax', 'ay', 'xsphx', 'xsphy', 'h', 'x0', 'y0', 'vx0', 'vy0', 'rho0', 'm', 'c']


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, lib, side, steps, kernel, fixed_dt, seq, q, sumdens=False, push=100.0, dt_rtol=0.0):
    os.environ["OSPH_LIB"] = lib                     # read by osph_b200.capi at import: this process binds the emulated build
    os.environ["OSPH_NCCL_LIB"] = os.path.join(os.path.dirname(lib), "libfake_nccl.so")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "offshore-sph_b200"), HERE):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from conftest import field_err
    from osph_b200 import capi, slabs, workloads as W
    assert capi.LIB_PATH == lib
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = W.dam_break_case(side, seed=11)
    f = case['pA']['label'] == 0
    case['pA']['vx'][f] += push                      # push the fluid across the slab faces
    pA, c = case['pA'], case['consts']
    if sumdens:
        c = dict(c, useSummationDensity=True)         # density from the kernel sum before every force evaluation
    cfg = capi.make_config(c, kernel, 'pec', capi.FP64, case['h'], reorder_every=3)
    with capi.Context(cfg) as single:
        single.upload(pA)
        single.step(steps, fixed_dt, 0.05)
        ref = single.download(pA.copy())
        ref_dt = single.dt_log()
    results = {}
    for chunk in (1, steps // 2):                    # one step per sequencer call (plain), then 7 per call (fused corrector)
        ctx = capi.Context(cfg)
        cuts, local_pA, ids = slabs.partition(pA, world, rank)
        if seq == 'nccl':      # the C++ loop of slab_nccl.cu against tests/emu/fake_nccl.cpp (OSPH_NCCL_LIB)
            run = slabs.NcclSlabRun(ctx, cuts, local_pA, ids, kernel, case['r0'], case['h'], torch.device('cpu'),
                                    mig_frac=0.2, ghost_frac=0.5, min_cap=256)
        elif seq == 'p2p':     # the C++ loop of slab_p2p.cu: IPC windows (POSIX shared memory here), mailbox kernels
            run = slabs.P2PSlabRun(ctx, cuts, local_pA, ids, kernel, case['r0'], case['h'], torch.device('cpu'),
                                   mig_frac=0.2, ghost_frac=0.5, min_cap=256)
        else:
            run = slabs.SlabRun(ctx, slabs.TorchComm(), cuts, local_pA, ids, kernel, case['r0'], case['h'], torch.device('cpu'),
                                mig_frac=0.2, ghost_frac=0.5, min_cap=256)
        moved, launches0 = 0, ctx.launch_count
        for k in range(0, steps, chunk):
            run.step(chunk, fixed_dt, 0.05)
            moved += sum(run.last_counts['mig_out'])
            if k + chunk == steps // 2:
                slabs.rebalance(run)                 # re-cut the slabs mid-run: results must not notice
        launches = ctx.launch_count - launches0
        got, seen = slabs.gather_global(run, pA, FIELDS)
        if hasattr(run, 'cadence_stats'):
            q.put((rank, 'cadence', chunk) + tuple(run.cadence_stats))
        dts = ctx.dt_log()
        status = ctx.sync()
        errs = {f_: field_err(got[f_], ref[f_]) for f_ in FIELDS}
        results[chunk] = (got, launches)
        if hasattr(run, 'close'):
            run.close()
        ctx.close()
        # dt: bit-identical where the Courant limit governs; where the force limit does, max |a|^2 carries the summation order
        dt_ok = bool(np.array_equal(dts, ref_dt)) or (dt_rtol > 0 and dts.shape == ref_dt.shape and bool(np.allclose(dts, ref_dt, rtol=dt_rtol, atol=0)))
        q.put((rank, chunk, errs, bool(np.all(seen == 1)), int(status), dt_ok, moved, len(pA)))
    a, b = results[1], results[steps // 2]
    same = all(np.array_equal(a[0][f_], b[0][f_]) for f_ in FIELDS)
    diff = max(field_err(a[0][f_], b[0][f_]) for f_ in FIELDS)
    q.put((rank, 'fused-vs-plain', same, a[1], b[1] - (1 if sumdens else 0), diff))   # (no corrector fusion with summation density)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kernel,fixed_dt,seq", [(3, 'cubic', FIXED_DT, 'python'),
                                                      (2, 'cubic', None, 'python'), (3, 'wendland', FIXED_DT, 'p2p'),
                                                      (2, 'cubic', None, 'p2p'), (3, 'cubic', None, 'nccl'),
                                                      (8, 'cubic', FIXED_DT, 'p2p')])
def test_emulated_slab_run_reproduces_single_rank_run(world, kernel, fixed_dt, seq):
    """fixed_dt None = the dynamic Courant / force time step, all-reduced every step (what bench.py --gpus N runs)."""
    _slab_case(emu_build.build(), world, kernel, fixed_dt, seq)


def test_emulated_slab_run_with_the_general_pair_kernel(monkeypatch):
    """The default build runs the uniform-h instantiation of the pair kernel in every slab test of this file (fixed h: ghost
    fluid particles carry the smoothing length their owner wrote, status stays 0 -- OSPH_S_H_NOT_UNIFORM would show).  Once with
    OSPH_UH=0: the general (pipelined) instantiation in slab mode."""
    monkeypatch.setenv("OSPH_UH", "0")
    _slab_case(emu_build.build(), 2, 'cubic', None, 'p2p')


def _slab_case(lib, world, kernel, fixed_dt, seq):
    import queue
    import time
    emu_build.build_fake_nccl()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lib, 60, 14, kernel, fixed_dt, seq, q)) for r in range(world)]
    [p.start() for p in procs]
    per_rank = 5 if seq == 'p2p' else 3            # the peer-memory sequencer also reports its cadence statistics (2 runs)
    res, t_end = [], time.time() + 400
    while len(res) < per_rank * world and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == per_rank * world and all(p.exitcode == 0 for p in procs)
    for r in res:
        if r[1] == 'cadence':
            continue
        if r[1] == 'fused-vs-plain':
            rank, _, same, launches_plain, launches_fused, diff = r
            # Bit-identical under the emulator's default (ascending) schedule.  Under OSPH_EMU_ORDER=reverse / shuffle the
            # pack kernel's atomically allocated record slots come out in another order -- as on the GPU -- and with them
            # the summation order over ghost neighbours: equal to rounding then.
            if os.environ.get("OSPH_EMU_ORDER", "").startswith(("reverse", "shuffle")):
                assert diff <= 1e-12, ("fused multi-step slab call differs from single-step calls", diff)
            else:
                assert same, ("fused multi-step slab call differs from single-step calls", diff)
            assert launches_fused < launches_plain          # the separate corrector passes are gone
            continue
        rank, chunk, errs, owned_once, status, dt_equal, moved, n = r
        assert owned_once, "a particle is owned by no rank or by two"
        assert status == 0
        assert dt_equal, "dt differs from the single-rank run"
        worst = max(errs, key=errs.get)
        assert errs[worst] <= 1e-11, (chunk, worst, errs[worst])
    if fixed_dt is not None:
        assert sum(r[6] for r in res if r[1] == 1) > 0, "no particle migrated: the test did not exercise the exchange"


@pytest.mark.parametrize("world,seq", [(2, 'python'), (3, 'p2p')])
def test_emulated_slab_run_with_summation_density(world, seq):
    """useSummationDensity in slab mode (reference src/Tools/SolverTools.py:129-140): the density pass also runs for the
    ghost particles, whose halo is twice as wide, so an owned particle reads the same neighbour densities as in the
    single-rank run."""
    import queue
    import time
    lib = emu_build.build()
    emu_build.build_fake_nccl()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lib, 40, 8, 'cubic', None, seq, q, True)) for r in range(world)]
    [p.start() for p in procs]
    per_rank = 5 if seq == 'p2p' else 3
    res, t_end = [], time.time() + 400
    while len(res) < per_rank * world and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == per_rank * world and all(p.exitcode == 0 for p in procs)
    for r in res:
        if r[1] == 'cadence':
            continue
        if r[1] == 'fused-vs-plain':
            assert r[5] <= 1e-12
            continue
        rank, chunk, errs, owned_once, status, dt_equal, moved, n = r
        assert owned_once and status == 0 and dt_equal
        worst = max(errs, key=errs.get)
        assert errs[worst] <= 1e-11, (chunk, worst, errs[worst])


def _worker_overflow(rank, world, port, lib, seq, q):
    os.environ["OSPH_LIB"] = lib
    os.environ["OSPH_NCCL_LIB"] = os.path.join(os.path.dirname(lib), "libfake_nccl.so")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "offshore-sph_b200"), HERE):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from osph_b200 import capi, slabs, workloads as W
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = W.dam_break_case(40, seed=2)
    pA, c = case['pA'], case['consts']
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    ctx = capi.Context(cfg)
    cuts, local_pA, ids = slabs.partition(pA, world, rank)
    cls = {'python': None, 'nccl': slabs.NcclSlabRun, 'p2p': slabs.P2PSlabRun}[seq]
    kw = dict(mig_frac=0.0, ghost_frac=0.0, min_cap=8)            # 8 halo records per face: far too few
    if cls is None:
        run = slabs.SlabRun(ctx, slabs.TorchComm(), cuts, local_pA, ids, 'cubic', case['r0'], case['h'], torch.device('cpu'), **kw)
    else:
        run = cls(ctx, cuts, local_pA, ids, 'cubic', case['r0'], case['h'], torch.device('cpu'), **kw)
    msg = ""
    try:
        run.step(2, None, 0.05)
    except Exception as e:      # noqa: BLE001
        msg = "%s: %s" % (type(e).__name__, e)
    q.put((rank, msg))
    dist.barrier()              # every rank got here: the failure was collective, nobody is left waiting in an exchange
    if hasattr(run, 'close'):
        run.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("seq", ['python', 'nccl', 'p2p'])
def test_exchange_region_overflow_is_a_loud_collective_error(seq):
    """Halo regions sized for 8 records: every rank must come back from the step with a capacity error that names the
    overflow (no silent truncation of the halo, no rank left spinning in a mailbox or a receive)."""
    import queue
    import time
    lib = emu_build.build()
    emu_build.build_fake_nccl()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker_overflow, args=(r, world, port, lib, seq, q)) for r in range(world)]
    [p.start() for p in procs]
    res, t_end = [], time.time() + 200
    while len(res) < world and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == world and all(p.exitcode == 0 for p in procs), res
    for rank, msg in res:
        assert "overflow" in msg.lower(), (rank, msg)


def _worker_silent_peer(rank, world, port, lib, q):
    os.environ["OSPH_LIB"] = lib
    os.environ["OSPH_P2P_SPIN_SECONDS"] = "3"
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "offshore-sph_b200"), HERE):
        sys.path.insert(0, p)
    import time
    import torch
    import torch.distributed as dist
    from osph_b200 import capi, slabs, workloads as W
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = W.dam_break_case(40, seed=2)
    pA, c = case['pA'], case['consts']
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    ctx = capi.Context(cfg)
    cuts, local_pA, ids = slabs.partition(pA, world, rank)
    run = slabs.P2PSlabRun(ctx, cuts, local_pA, ids, 'cubic', case['r0'], case['h'], torch.device('cpu'),
                           mig_frac=0.2, ghost_frac=0.5, min_cap=256)
    msg, t0 = "", time.time()
    if rank == 0:
        msg = "stayed out"                      # e.g. a Python exception on this rank before its step: it never joins the exchange
    else:
        try:
            run.step(2, None, 0.05)
        except Exception as e:      # noqa: BLE001
            msg = "%s: %s" % (type(e).__name__, e)
    q.put((rank, msg, time.time() - t0))
    dist.barrier()
    run.close()
    ctx.close()
    dist.destroy_process_group()


def test_a_silent_peer_is_a_timeout_not_a_hang():
    """Peer-memory sequencer: a rank that never joins the step (it failed elsewhere) must not leave the others inside a
    mailbox kernel for ever.  The waits are bounded (OSPH_P2P_SPIN_SECONDS): the waiting rank comes back with OSPH_E_PEER."""
    import queue
    import time
    lib = emu_build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_silent_peer, args=(r, 2, port, lib, q)) for r in range(2)]
    [p.start() for p in procs]
    res, t_end = [], time.time() + 120
    while len(res) < 2 and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == 2 and all(p.exitcode == 0 for p in procs), res
    waited = [r for r in res if r[0] == 1][0]
    assert "did not answer" in waited[1] or "peer" in waited[1].lower(), waited
    assert waited[2] < 60, waited


@pytest.mark.parametrize("world,fixed_dt", [(2, None), (3, 2e-4)])
def test_emulated_slab_cadence_reuses_the_halo_lists(world, fixed_dt):
    """Fine-cell regime (pair radius < reference cell), peer-memory sequencer: between two sorting steps the ranks re-send
    the SAME halo particles into the SAME record slots, migration waits for the next sort and nobody sorts -- decided by all
    ranks alike from the all-gathered displacement.  Must be invisible: the gathered result equals the single-rank run,
    every particle owned once, dt identical, status 0 (OSPH_S_SKIN_EXHAUSTED would show here), and steps were reused."""
    import queue
    import time
    lib = emu_build.build()
    emu_build.build_fake_nccl()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    steps = 16
    procs = [ctx.Process(target=_worker, args=(r, world, port, lib, 100, steps, 'cubic', fixed_dt, 'p2p', q, False, 25.0, 1e-11)) for r in range(world)]
    [p.start() for p in procs]
    res, t_end = [], time.time() + 900
    while len(res) < 5 * world and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(60) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == 5 * world and all(p.exitcode == 0 for p in procs), res
    reused = 0
    for r in res:
        if r[1] == 'cadence':
            rank, _, chunk, sorts, reuses = r
            assert sorts + reuses == steps and sorts >= 2
            reused += reuses
            continue
        if r[1] == 'fused-vs-plain':
            assert r[5] <= 1e-11
            continue
        rank, chunk, errs, owned_once, status, dt_equal, moved, n = r
        assert owned_once and status == 0 and dt_equal, r
        worst = max(errs, key=errs.get)
        assert errs[worst] <= 1e-11, (chunk, worst, errs[worst])
    print("cadence:", [r[2:] for r in res if r[1] == 'cadence'])
    assert reused > 0, "no step reused the binning: the test did not exercise the cadence"


def _worker_tank(rank, world, port, lib, q):
    os.environ["OSPH_LIB"] = lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "offshore-sph_b200"), HERE):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from conftest import field_err
    from osph_b200 import capi, slabs, workloads as W
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = W.tank_case(50, h=None, useXSPH=False, seed=3)            # Containment set-up: h = 1.3 sqrt(m / rho), refreshed every step
    pA, c = case['pA'], case['consts']
    pA['vx'][pA['label'] == 0] += 0.5
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, None)
    steps = 20
    with capi.Context(cfg) as single:
        single.upload(pA); single.step(steps, None, 0.0)
        ref, ref_dt = single.download(pA.copy()), single.dt_log()
    ctx = capi.Context(cfg)
    cuts, local_pA, ids = slabs.partition(pA, world, rank)
    run = slabs.P2PSlabRun(ctx, cuts, local_pA, ids, 'cubic', case['r0'], float(pA['h'].max()), torch.device('cpu'),
                           mig_frac=0.2, ghost_frac=0.6, min_cap=256)
    run.step(steps, None, 0.0)
    fields = ['x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay', 'drho', 'h']
    got, seen = slabs.gather_global(run, pA, fields)
    errs = {k: field_err(got[k], ref[k]) for k in fields}
    q.put((rank, run.cadence_stats, ctx.sync(), bool(np.all(seen == 1)), max(errs.values()), max(errs, key=errs.get),
           bool(np.allclose(ctx.dt_log(), ref_dt, rtol=1e-10, atol=0))))
    dist.barrier()
    run.close(); ctx.close()
    dist.destroy_process_group()


def test_emulated_slab_cadence_with_dynamic_h():
    """The Containment set-up (smoothing length refreshed from the density every step, so the pair radius moves) through the
    slab cadence: the host plans the reuse with a margin on the radius, the device verifies it (status must stay 0)."""
    import queue
    import time
    lib = emu_build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_tank, args=(r, 2, port, lib, q)) for r in range(2)]
    [p.start() for p in procs]
    res, t_end = [], time.time() + 400
    while len(res) < 2 and time.time() < t_end:
        try:
            res.append(q.get(timeout=1.0))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    [p.join(30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) == 2 and all(p.exitcode == 0 for p in procs), res
    for rank, (sorts, reuses), status, owned_once, err, worst, dt_ok in res:
        assert sorts + reuses == 20 and reuses >= 10, (sorts, reuses)
        assert status == 0 and owned_once and dt_ok
        assert err <= 1e-11, (worst, err)
