"""Explicit Euler; interface as reference src/Integrators/Euler.py:7-26 (predict is the identity)."""
from src.Integrators.Integrator import Integrator


class Euler(Integrator):
    osph_name = 'euler'

    def predict(self, dt, pA, damping: float = 0.0):
        return pA
