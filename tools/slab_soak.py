"""torchrun script: the slab cadence over a long run with real motion -- the collapsing dam-break column (gate removed,
dynamic dt), peer-memory sequencer, against the same run on one GPU.  Particles cross the slab faces for good, so the run
goes through many cycles of sort (migration, fresh halo lists) and reuse.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tools/slab_soak.py [side=400] [steps=1500]
"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from osph_b200 import capi, slabs, workloads as W      # noqa: E402
from conftest import field_err                          # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 400
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
case = W.dam_break_case(side, seed=3, temp_wall=False)
pA, c = case['pA'], case['consts']
cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'], device=local)
FIELDS = ['x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay', 'drho']
with capi.Context(cfg) as single:
    single.upload(pA)
    single.step(steps, None, 0.0)
    ref = single.download(pA.copy())
    ref_dt = single.dt_log()
ctx = capi.Context(cfg)
torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=local))
cuts, local_pA, ids = slabs.partition(pA, world, rank)
run = slabs.P2PSlabRun(ctx, cuts, local_pA, ids, 'cubic', case['r0'], case['h'], torch.device('cuda', local),
                       mig_frac=0.1, ghost_frac=0.5, min_cap=4096)
owned0 = ctx.num_active
done = 0
while done < steps:
    k = min(100, steps - done)
    run.step(k, None, 0.0)
    done += k
got, seen = slabs.gather_global(run, pA, FIELDS)
dts = ctx.dt_log()
status = ctx.sync()
errs = {f: field_err(got[f], ref[f]) for f in FIELDS}
sorts, reuses = run.cadence_stats
owned = torch.tensor([ctx.num_active - owned0], device='cuda', dtype=torch.float64)
allowned = torch.zeros(world, device='cuda', dtype=torch.float64)
dist.all_gather_into_tensor(allowned, owned)
ok = bool(np.all(seen == 1)) and status == 0 and max(errs.values()) <= 1e-7 and np.allclose(dts, ref_dt, rtol=1e-8, atol=0)
if rank == 0:
    print(json.dumps(dict(side=side, particles=len(pA), ranks=world, steps=steps, t=float(ref_dt[:, 0].sum()), sorting_steps=sorts,
                          reusing_steps=reuses, status=status, every_particle_owned_once=bool(np.all(seen == 1)),
                          net_particles_gained_per_rank=[int(v) for v in allowned.cpu().numpy()],
                          worst_field_error_vs_one_gpu=max(errs.values()), worst_field=max(errs, key=errs.get),
                          dt_max_rel_diff=float(np.max(np.abs(dts[:, 0] / ref_dt[:, 0] - 1.0))), ok=ok)), flush=True)
flag = torch.tensor([0 if ok else 1], device='cuda'); dist.all_reduce(flag)
torch.cuda.synchronize()
torch.cuda.set_stream(torch.cuda.default_stream())
run.close(); ctx.close()
dist.barrier(); dist.destroy_process_group()
sys.exit(1 if flag.item() else 0)
