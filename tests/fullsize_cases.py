"""Bodies of tests/test_gpu_fullsize.py, which runs each of them in a child process (a CUDA fault at this size
must fail one test, not leave a sticky error that blanks the rest of the GPU suite).

GPU, BASELINE configs[1] size (dam break N = 1000, 1 009 603 particles): properties that do not need the
full CPU oracle (≈1 min per step there), plus the oracle on a strided sample of the particles.

The reference search visits 9 x 1600 candidates per particle at this size, so the oracle evaluates every 997th
fluid particle (~1000 of them) against ALL particles; the GPU must agree on those to 1e-10.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "offshore-sph_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from conftest import field_err
from oracle import oracle as O
from osph_b200 import capi
from osph_b200 import workloads as W

N_SIDE = 1000
STRIDE = 997


def make_case():
    return W.dam_break_case(N_SIDE, seed=0)


def sampled_oracle_parity_at_full_size(kernel):
    case = make_case()
    pA, c = case['pA'], case['consts']
    P = O.Particles.from_aos(pA)
    w = O.wcsph(c['height'], c['r0'], c['rho0'], True)
    grid = O.Grid(P)
    O.loop(P, w, grid, kernel, STRIDE, 0)
    sample = np.flatnonzero((np.arange(P.n) % STRIDE == 0) & (P.label == 0))
    assert len(sample) > 900
    cfg = capi.make_config(c, kernel, 'pec', capi.FP64, case['h'], keep_h=True)
    with capi.Context(cfg) as ctx:
        ctx.upload(pA)
        ctx.compute()
        g, cells = ctx.cells()
        assert np.array_equal(cells, grid.cell_ids())                 # every one of the 1 M cell assignments
        out = ctx.download(pA.copy())
        assert ctx.sync() == 0
    for f in ('drho', 'ax', 'ay', 'xsphx', 'xsphy'):
        ref = getattr(P, f)[sample]
        scale = np.maximum(np.abs(ref), np.abs(ref).max())
        assert np.max(np.abs(out[f][sample] - ref) / scale) <= 1e-10, f
    assert field_err(out['p'], P.p) <= 1e-12                          # EOS runs for all particles in the oracle


def full_size_invariants():
    """Determinism, row-order invariance, and agreement of the two precisions at 1 M particles."""
    case = make_case()
    pA, c = case['pA'], case['consts']
    cfg = capi.make_config(c, 'cubic', 'pec', capi.FP64, case['h'])
    with capi.Context(cfg) as ctx:
        ctx.upload(pA); ctx.step(3, None, 0.05); a = ctx.download(pA.copy()); dts = ctx.dt_log()
        ctx.upload(pA); ctx.step(3, None, 0.05); b = ctx.download(pA.copy())
        assert np.array_equal(ctx.dt_log(), dts)
        assert a.tobytes() == b.tobytes()                             # bit-reproducible run to run
        # shuffle the host rows: every particle must get the same answer up to summation order
        perm = np.random.default_rng(1).permutation(len(pA))
        ctx.upload(np.ascontiguousarray(pA[perm])); ctx.step(3, None, 0.05)
        s = ctx.download(np.ascontiguousarray(pA[perm]).copy())
        assert np.allclose(ctx.dt_log(), dts, rtol=1e-13, atol=0)
        for f in ('x', 'y', 'vx', 'vy', 'rho', 'ax', 'ay', 'drho', 'p'):
            assert field_err(s[f], a[f][perm]) <= 1e-11, f
        # neighbour relation is symmetric for equal h: every (i, j) has its (j, i)
        off, idx = ctx.neighbours_csr()
        deg = np.diff(off)
        fluid = s['label'] == 0
        assert deg[~fluid].sum() == 0 and 60 < deg[fluid].mean() < 80  # ~72 = pi 4.8^2 within q <= 3
        src = np.repeat(np.arange(len(deg)), deg)
        ff = fluid[idx]                                               # fluid-fluid pairs only (walls have no list)
        a_key = src[ff].astype(np.int64) * len(deg) + idx[ff]
        b_key = idx[ff].astype(np.int64) * len(deg) + src[ff]
        assert np.array_equal(np.sort(a_key), np.sort(b_key))
    cfg32 = capi.make_config(c, 'cubic', 'pec', capi.FP32, case['h'])
    with capi.Context(cfg32) as ctx:
        ctx.upload(pA); ctx.step(3, None, 0.05); f32 = ctx.download(pA.copy())
    assert np.max(np.hypot(f32['x'] - a['x'], f32['y'] - a['y'])) < 1e-5 * case['r0']
    assert field_err(f32['rho'], a['rho']) < 1e-6


if __name__ == "__main__":
    if os.environ.get("OSPH_EMU") == "1":
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build as emu_build
        capi.LIB_PATH, capi._lib = emu_build.build(), None
    name, args = sys.argv[1], sys.argv[2:]
    globals()[name](*args)
    print("ok", name, *args)
