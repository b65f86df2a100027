"""Cost of moving the Coupled rows of a structure around one host-side integrator call: whole-array round trip
(reference src/Solver.py:381-398 shape: download all, edit rows, upload all) against osph_download_rows /
osph_upload_rows.      python tools/coupling_rows.py [N]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))
from osph_b200 import capi, workloads as W  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
case = W.tank_case(N, h=1.3 / N, useXSPH=True, seed=2, coupled_row=True)
pA = case['pA']
rows = np.flatnonzero(pA['label'] == 3).astype(np.int64)
ctx = capi.Context(capi.make_config(case['consts'], 'wendland', 'pec', capi.FP64, case['h']))
ctx.upload(pA)
ctx.step(5)
ctx.sync()
reps = 10


def timed(fn):
    fn()
    ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


def whole():
    ctx.download(pA)
    pA['y'][rows] += 0.0
    ctx.upload(pA)
    ctx.step(1)


def few():
    ctx.download_rows(rows, pA)
    pA['y'][rows] += 0.0
    ctx.upload_rows(rows, pA)
    ctx.step(1)


print("particles %d, coupled rows %d: ms per (round trip + step): whole array %.2f | coupled rows only %.2f | step alone %.2f"
      % (len(pA), len(rows), timed(whole), timed(few), timed(lambda: ctx.step(1))))
ctx.close()
