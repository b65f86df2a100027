"""torchrun script (one rank per GPU): slab-decomposed steps must reproduce the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--n 120] [--steps 12]

Every rank also runs the whole problem alone on its own GPU; the gathered slab result must agree with it to
summation order (FP64: 1e-11 norm-wise), every particle must be owned exactly once, and dt must be identical.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))

from osph_b200 import capi, slabs, workloads as W      # noqa: E402
from conftest import field_err                          # noqa: E402

FIXED_DT = 2e-4     # the force criterion (a^2!) would shrink a dynamic dt until nothing crosses a slab face
FIELDS = ['x', 'y', 'vx', 'vy', 'rho', 'p', 'drho', 'ax', 'ay', 'xsphx', 'xsphy', 'h', 'x0', 'y0', 'vx0', 'vy0', 'rho0', 'm', 'c']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--side', type=int, default=120)
    ap.add_argument('--steps', type=int, default=30)
    a = ap.parse_args()
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok = True
    for kernel, prec, tol, seq in (('wendland', capi.FP64, 1e-11, 'nccl'), ('wendland', capi.FP64, 1e-11, 'python'),
                                   ('wendland', capi.FP64, 1e-11, 'p2p'), ('cubic', capi.FP64, 1e-11, 'p2p'),
                                   ('cubic', capi.FP64, 1e-11, 'nccl'), ('gaussian', capi.FP32, 2e-3, 'nccl')):
        case = W.dam_break_case(a.side, seed=11)
        # give the fluid a push towards +x so particles cross the slab faces during the run
        f = case['pA']['label'] == 0
        case['pA']['vx'][f] += 100.0
        pA, c = case['pA'], case['consts']
        cfg = capi.make_config(c, kernel, 'pec', prec, case['h'], device=local, reorder_every=3)
        with capi.Context(cfg) as single:
            single.upload(pA)
            single.step(a.steps, FIXED_DT, 0.05)
            ref = single.download(pA.copy())
            ref_dt = single.dt_log()
        ctx = capi.Context(cfg)
        torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=local))
        cuts, local_pA, ids = slabs.partition(pA, world, rank)
        if seq == 'python':
            run = slabs.SlabRun(ctx, slabs.TorchComm(), cuts, local_pA, ids, kernel, case['r0'], case['h'],
                                torch.device('cuda', local))
        elif seq == 'p2p':
            # capacities derived from the (rank-dependent) local counts: the ranks must agree on one window layout
            run = slabs.P2PSlabRun(ctx, cuts, local_pA, ids, kernel, case['r0'], case['h'], torch.device('cuda', local),
                                   mig_frac=0.2, ghost_frac=0.5, min_cap=256)
        else:
            run = slabs.NcclSlabRun(ctx, cuts, local_pA, ids, kernel, case['r0'], case['h'], torch.device('cuda', local))
        moved = 0
        for k in range(a.steps):
            run.step(1, FIXED_DT, 0.05)
            moved += sum(run.last_counts['mig_out'])
            if k == a.steps // 2:
                slabs.rebalance(run)                   # re-cut the slabs mid-run: results must not notice
        got, seen = slabs.gather_global(run, pA, FIELDS)
        dts = ctx.dt_log()
        status = ctx.sync()
        errs = {f_: field_err(got[f_], ref[f_]) for f_ in FIELDS}
        worst = max(errs.values())
        tot_moved = torch.tensor([moved], device='cuda'); dist.all_reduce(tot_moved)
        good = bool(np.all(seen == 1)) and worst <= tol and status == 0 and \
            np.allclose(dts, ref_dt, rtol=1e-12 if prec == capi.FP64 else 1e-4, atol=0)
        cadence = " sorted/reused=%d/%d" % run.cadence_stats if hasattr(run, 'cadence_stats') else ""
        if rank == 0:
            print("%-9s %s %-6s ranks=%d n=%d steps=%d migrants=%d worst_err=%.2e (%s) dt_equal=%s%s -> %s" % (
                kernel, 'fp64' if prec == capi.FP64 else 'fp32', seq, world, len(pA), a.steps, int(tot_moved.item()), worst,
                max(errs, key=errs.get), np.allclose(dts, ref_dt, rtol=1e-12, atol=0), cadence, 'OK' if good else 'FAIL'), flush=True)
        ok = ok and good and (a.steps < 10 or int(tot_moved.item()) > 0)
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.default_stream())      # never leave torch on a stream about to be destroyed
        if hasattr(run, 'close'):
            run.close()
        ctx.close()
    flag = torch.tensor([0 if ok else 1], device='cuda'); dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == '__main__':
    main()
