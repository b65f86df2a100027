"""XSPH velocity correction on a table of computed neighbours, interface of reference src/Equations/XSPH.py:6-31
(sums over every label; boundary particles carry m = 0).  Fused into csrc/pair.cu inside Solver.run(); stand-alone
calls run on the device through osph_leaf_equations."""
from typing import List

from osph_b200 import capi


def XSPH(epsilon, p, comp) -> List[float]:
    return capi.leaf_equations(p, comp, epsilon=epsilon)['xsph']
