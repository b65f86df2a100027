"""Run one force evaluation of a golden case (for compute-sanitizer / racecheck runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from osph_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else 'dambreak20_wendland'
prec = capi.FP32 if len(sys.argv) > 2 and sys.argv[2] == 'fp32' else capi.FP64
g, meta, pA = load_golden(name)
cfg = capi.make_config(meta['consts'] | {'useXSPH': meta['useXSPH']}, meta['kernel'], 'pec', prec, meta['fixed_h'], keep_h=True)
with capi.Context(cfg) as ctx:
    ctx.upload(pA)
    ctx.compute()
    ctx.step(2, None, 0.05)
    out = ctx.download(pA.copy())
    print(name, 'ok', float(out['ax'].sum()))
