"""Per-step cost of exporting x, y, p (the Solver default) at bench size: blocking osph_download_fields + host merge
(the shape of reference src/Solver.py:477-486) against the double-buffered row-space export.
    python tools/export_overlap.py [N] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))
from osph_b200 import capi, workloads as W  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
case = W.dam_break_case(N)
pA = case['pA']
ctx = capi.Context(capi.make_config(case['consts'], 'cubic', 'pec', capi.FP64, case['h']))
ctx.upload(pA)
ctx.step(10, damping=0.05)
ctx.sync()
names = ['x', 'y', 'p']
idx = ~pA['deleted']


def timed(fn):
    ctx.sync()
    t = time.perf_counter()
    out = fn()
    ctx.sync()
    return (time.perf_counter() - t) / steps * 1e3, out


def plain():
    for _ in range(steps):
        ctx.step(1, damping=0.05)


def blocking():
    keep = []
    for _ in range(steps):
        ctx.step(1, damping=0.05)
        cols = ctx.download_fields(names)
        for k in names:                                  # the reference's merge into full-length arrays
            full = np.copy(pA[k])
            full[idx] = cols[k]
            keep.append(full)
    return keep


def overlapped():
    keep, pending = [], []
    for _ in range(steps):
        ctx.step(1, damping=0.05)
        pending.append(ctx.export_begin(names, rows=True))
        if len(pending) > 1:
            keep.extend(ctx.export_end(pending.pop(0)).values())
    while pending:
        keep.extend(ctx.export_end(pending.pop(0)).values())
    return keep


t0, _ = timed(plain)
t1, a = timed(blocking)
t2, b = timed(overlapped)
print("particles %d  steps %d  ms/step: no export %.3f | blocking export + host merge %.3f | async row-space export %.3f"
      % (len(pA), steps, t0, t1, t2))
assert len(a) == len(b)
ctx.close()
