"""Kinetic energy of the rows passed in (reference src/Equations/KineticEnergy.py:6-12).
The reference reads J rows of an array that may be shorter (src/Solver.py:416); here the rows of pA count."""
import numpy as np


def KineticEnergy(J, pA) -> float:
    from src.Equations.TimeStep import _scratch_context
    from src.Common import particle_dtype
    arr = np.ascontiguousarray(pA).astype(particle_dtype, copy=True)
    arr['label'] = 0
    ctx = _scratch_context(arr)
    try:
        return ctx.kinetic_energy()
    finally:
        ctx.close()
