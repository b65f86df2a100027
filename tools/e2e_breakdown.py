"""Where the end-to-end step (upload + step + download with host buffers) spends its time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))
import torch
from osph_b200 import capi, workloads as W
case = W.dam_break_case(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
pA = case['pA']; n = len(pA)
ctx = capi.Context(capi.make_config(case['consts'], 'cubic', 'pec', capi.FP64, case['h']))
hbuf = torch.empty(n * 154, dtype=torch.uint8, pin_memory=True); host = hbuf.numpy().view(pA.dtype); host[:] = pA
ctx.upload(host); ctx.step(3, None, 0.05); ctx.download(host)
T = [0, 0, 0]
for _ in range(10):
    t0 = time.perf_counter(); ctx.upload(host); ctx.sync()
    t1 = time.perf_counter(); ctx.step(1, None, 0.05); ctx.sync()
    t2 = time.perf_counter(); ctx.download(host)
    t3 = time.perf_counter()
    T[0] += t1 - t0; T[1] += t2 - t1; T[2] += t3 - t2
print("n=%d upload %.2f ms  step %.2f ms  download %.2f ms  (device phase timers: %s)" % (
    n, T[0] * 100, T[1] * 100, T[2] * 100, {k: round(v, 4) for k, v in ctx.timers().items()}))
