"""Shepard filter of a neighbour weight array (reference src/Equations/Shepard.py:5-12).
Host helper for user coupling callbacks."""
import numpy as np

from src.Common import ParticleType


def Shepard(w: np.array, labels: np.array, m_b: np.array, rho_b: np.array):
    w = np.asarray(w, dtype=np.float64)
    ok = (np.asarray(rho_b) >= 1e-3) & (np.asarray(labels) == ParticleType.Fluid)
    w_tilde = np.sum(w[ok] * np.asarray(m_b)[ok] / np.asarray(rho_b)[ok])
    return w / w_tilde
