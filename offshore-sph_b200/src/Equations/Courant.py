"""Courant condition alpha * min(h) / max(c), interface of reference src/Equations/Courant.py:4-31 (the Solver itself
uses TimeStep).  Runs on the device through osph_leaf_courant."""
from osph_b200 import capi


def Courant(alpha, h, c) -> float:
    assert len(h) == len(c)
    return capi.leaf_courant(alpha, h, c)
