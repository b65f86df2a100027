"""Inputs of the leaf-equation golden vectors (tests/golden/equations_leaf.npz), shared by the generator
(make_golden.py, which feeds them to the UNMODIFIED reference equations) and by tests/test_oracle_golden.py (which feeds
them to the oracle).  The first three are the synthetic tables of the reference's own equation tests:

  momentum_reftest     reference test/test_numba_momentum.py:21-66   (10 000 neighbours, linspace columns)
  continuity_reftest   reference test/test_numba_continuity.py:12-29 (10 000 000 there; 100 000 here, same construction)
  boundary_reftest     reference test/test_eq_boundary.py:9-26       (one wall particle left of a fluid particle)
  random               64 neighbours with mixed labels, approaching and separating pairs, h and rho spread
"""
import numpy as np

# field names and order of the reference's computed_dtype (src/Common.py:59-105)
COMP_FIELDS = ['m', 'p', 'rho', 'h', 'q', 'c', 'r', 'w', 'dw_x', 'dw_y', 'x', 'y', 'vx', 'vy']
COMP_DTYPE = np.dtype([('label', np.int8)] + [(f, np.float64) for f in COMP_FIELDS])


def momentum_reftest(num=10_000):
    lin = np.linspace(0, 2000, num)
    comp = np.zeros(num, dtype=COMP_DTYPE)
    comp['p'] = np.linspace(0, 10_000, num)
    comp['rho'] = 1025.0
    comp['h'] = 1.3
    comp['x'] = lin; comp['y'] = lin
    comp['r'] = np.sqrt(2.0) * lin                   # cdist(xij, xij)[0, :] of the reference test
    comp['dw_x'] = lin; comp['dw_y'] = lin
    comp['vx'] = lin; comp['vy'] = lin
    comp['m'] = 1.0
    cs = 10.0 * np.sqrt(2 * 9.81 * 1.0)
    return dict(rho=1000.0, c=cs, h=1.3, m=1.0, p=0.0), comp


def continuity_reftest(num=100_000):
    lin = np.linspace(0, 2000, num)
    comp = np.zeros(num, dtype=COMP_DTYPE)
    comp['m'] = 1.0
    comp['vx'] = lin; comp['vy'] = lin; comp['dw_x'] = lin; comp['dw_y'] = lin
    return comp


def boundary_reftest():
    """A wall particle at (-0.5, 0) seen from a fluid particle at the origin, r0 = 1 (test_eq_boundary.py)."""
    comp = np.zeros(2, dtype=COMP_DTYPE)
    comp['label'] = [1, 0]                           # the wall, and a fluid neighbour that must be ignored
    comp['x'] = [0.5, -0.3]; comp['y'] = [0.0, 0.1]
    comp['r'] = np.hypot(comp['x'], comp['y'])
    return comp


def random_table(J=64, seed=3):
    rng = np.random.default_rng(seed)
    comp = np.zeros(J, dtype=COMP_DTYPE)
    comp['label'] = rng.choice([0, 0, 0, 1, 3], size=J)
    comp['x'] = rng.uniform(-0.2, 0.2, J); comp['y'] = rng.uniform(-0.2, 0.2, J)
    comp['r'] = np.hypot(comp['x'], comp['y'])
    comp['r'][5] = 0.0; comp['x'][5] = 0.0; comp['y'][5] = 0.0          # coincident neighbour (r <= 1e-12 guard)
    comp['label'][5] = 1
    comp['vx'] = rng.normal(size=J) * 2.0; comp['vy'] = rng.normal(size=J) * 2.0
    comp['rho'] = 1000.0 + rng.normal(size=J) * 15.0
    comp['p'] = rng.uniform(0.0, 3.0e4, J)
    comp['h'] = rng.uniform(0.09, 0.13, J)
    comp['c'] = rng.uniform(0.0, 40.0, J)
    comp['m'] = np.where(comp['label'] == 1, 0.0, rng.uniform(0.5, 1.5, J))
    comp['w'] = rng.uniform(0.0, 30.0, J)
    comp['dw_x'] = rng.normal(size=J) * 50.0; comp['dw_y'] = rng.normal(size=J) * 50.0
    comp['q'] = comp['r'] / comp['h']
    return dict(rho=1003.0, c=44.3, h=0.11, m=1.0, p=1.7e4), comp
