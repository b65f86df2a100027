import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))
from osph_b200 import capi, workloads as W
side = int(sys.argv[1]) if len(sys.argv) > 1 else 160
case = W.dam_break_case(side, seed=11)
f = case['pA']['label'] == 0
case['pA']['vx'][f] += 100.0
pA, c = case['pA'], case['consts']
cfg = capi.make_config(c, 'wendland', 'pec', capi.FP64, case['h'], reorder_every=3)
with capi.Context(cfg) as ctx:
    ctx.upload(pA)
    for k in range(30):
        ctx.step(1, 2e-4, 0.05)
        st = ctx.sync()
        print(k, 'status', st, flush=True)
    out = ctx.download(pA.copy())
print('ok', float(out['x'].max()))
