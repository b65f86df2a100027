"""Summation density of ONE particle from its neighbour arrays (reference src/Equations/SummationDensity.py:6-13).
Host helper for user coupling callbacks (IceBreak pressure probe: a few dozen neighbours per call)."""
import numpy as np

from src.Common import ParticleType


def SummationDensity(labels: np.array, m: np.array, w: np.array) -> float:
    f = np.asarray(labels) == ParticleType.Fluid
    return float(np.sum(np.asarray(m)[f] * np.asarray(w)[f]))
