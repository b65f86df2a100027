"""Lennard-Jones boundary force (Monaghan 1992) on a table of computed neighbours, interface of reference
src/Equations/BoundaryForce.py:7-42: non-fluid neighbours with 1e-12 < r <= r0.  Fused into csrc/pair.cu inside
Solver.run(); stand-alone calls run on the device through osph_leaf_equations."""
from typing import List

from osph_b200 import capi


def BoundaryForce(r0, D, p1, p2, p, comp) -> List[float]:
    return capi.leaf_equations(p, comp, r0=r0, D=D, p1=p1, p2=p2)['f']
