"""Sort cadence over a long run at the bench size: the collapsing dam-break column (gate removed, dynamic dt), FP64, run
twice -- binning reused between sorts (default) and sorted at every build (OSPH_SKIN=0) -- and compared at checkpoints.
The two runs may differ by summation order only.

    python tools/cadence_soak.py [N=1000] [steps=3000]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))
import numpy as np
from osph_b200 import capi, workloads as W

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
case = W.dam_break_case(N, seed=0, temp_wall=False)
pA, c = case['pA'], case['consts']
ctxs = {}
for skin in ('auto', '0'):
    os.environ['OSPH_SKIN'] = skin
    ctxs[skin] = capi.Context(capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h']))
    ctxs[skin].upload(pA)
fluid = pA['label'] == 0
done, rows = 0, []
for chunk in (10, 90, 400, 500, 1000, 1000, 2000, 5000):
    if done >= steps:
        break
    chunk = min(chunk, steps - done)
    out = {}
    for skin, ctx in ctxs.items():
        ctx.step(chunk, None, 0.0)
        out[skin] = ctx.download(pA.copy())
    done += chunk
    a, b = out['0'], out['auto']
    err = lambda f: float(np.max(np.abs(a[f] - b[f])[fluid]) / max(np.max(np.abs(a[f][fluid])), 1e-300))
    builds, sorts = ctxs['auto'].sort_stats()
    rows.append(dict(steps=done, t=float(ctxs['0'].dt_log()[:, 0].sum()) if done == chunk else None,
                     builds=builds, sorts=sorts, status=[ctxs['auto'].sync(), ctxs['0'].sync()],
                     **{f: err(f) for f in ('x', 'y', 'vx', 'vy', 'rho', 'p', 'ax', 'ay', 'drho')}))
    print(json.dumps(rows[-1]))
off_a, idx_a = ctxs['auto'].neighbours_csr()
off_b, idx_b = ctxs['0'].neighbours_csr()
same = bool(np.array_equal(off_a, off_b) and np.array_equal(idx_a, idx_b))
print(json.dumps(dict(N=N, particles=len(pA), steps=done, neighbour_lists_identical_at_end=same, rows=rows)))
