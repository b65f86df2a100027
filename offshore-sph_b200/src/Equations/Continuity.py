"""Continuity equation on a table of computed neighbours, interface of reference src/Equations/Continuity.py:5-17.
Inside Solver.run() the sum is part of the fused pair kernel (csrc/pair.cu); this function serves stand-alone
callers (WCSPH.compute_density_change, the reference's test/test_numba_continuity.py) and runs on the device
through osph_leaf_equations."""
from osph_b200 import capi


def Continuity(p, comp) -> float:
    return capi.leaf_equations(p, comp)['drho']
