"""Drift-kick-drift Verlet; constructor as reference src/Integrators/Verlet.py:10-24."""
from src.Integrators.Integrator import Integrator


class Verlet(Integrator):
    osph_name = 'verlet'

    def __init__(self, useXSPH: bool = True):
        self.useXSPH = bool(useXSPH)
