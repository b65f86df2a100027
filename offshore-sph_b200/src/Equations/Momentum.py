"""Monaghan momentum equation with artificial viscosity on a table of computed neighbours, interface of reference
src/Equations/Momentum.py:6-57 (gravity is added by WCSPH.compute_acceleration).  Fused into csrc/pair.cu inside
Solver.run(); stand-alone calls run on the device through osph_leaf_equations."""
from typing import List

from osph_b200 import capi


def Momentum(alpha, beta, p, comp) -> List[float]:
    return capi.leaf_equations(p, comp, alpha=alpha, beta=beta)['a']
