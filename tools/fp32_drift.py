"""FP32 performance mode vs FP64 validation mode over a dam-break run (north star: "drift over a dam-break run
bounded and reported for FP32").  Same initial state, same number of steps, dynamic dt in both runs.

    python tools/fp32_drift.py [N=200] [steps=2000]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))
import numpy as np
from osph_b200 import capi, workloads as W

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
case = W.dam_break_case(N, seed=0, temp_wall=False)          # gate already removed: the column collapses
pA, c, r0 = case['pA'], case['consts'], case['r0']
ctxs = {p: capi.Context(capi.make_config(c, 'wendland', 'pec', p, case['h'])) for p in (capi.FP64, capi.FP32)}
for ctx in ctxs.values():
    ctx.upload(pA)
fluid = pA['label'] == 0
rows, done, t = [], 0, {capi.FP64: 0.0, capi.FP32: 0.0}
for chunk in (1, 9, 90, 400, 500, 1000, 2000, 4000):
    if done >= steps:
        break
    chunk = min(chunk, steps - done)
    out = {}
    for p, ctx in ctxs.items():
        ctx.step(chunk, None, 0.0)
        out[p] = ctx.download(pA.copy())
        t[p] += float(ctx.dt_log()[:, 0].sum())
    done += chunk
    a, b = out[capi.FP64], out[capi.FP32]
    d = np.hypot(a['x'] - b['x'], a['y'] - b['y'])[fluid]
    ke = lambda o: float(np.sum(0.5 * o['m'] * (o['vx'] ** 2 + o['vy'] ** 2)))
    rows.append(dict(steps=done, t_fp64=t[capi.FP64], t_fp32=t[capi.FP32],
                     pos_max_over_r0=float(d.max() / r0), pos_rms_over_r0=float(np.sqrt(np.mean(d ** 2)) / r0),
                     rho_max_rel=float(np.max(np.abs(a['rho'] - b['rho'])[fluid] / a['rho'][fluid])),
                     ke_rel=abs(ke(a) - ke(b)) / max(ke(a), 1e-300),
                     front_x_fp64=float(a['x'][fluid].max()), front_x_fp32=float(b['x'][fluid].max())))
    print(json.dumps(rows[-1]))
print(json.dumps(dict(N=N, particles=len(pA), r0=r0, kernel='wendland', rows=rows)))
