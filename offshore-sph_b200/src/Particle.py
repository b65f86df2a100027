"""Single-particle convenience record with the attributes of reference src/Particle.py:4-39 (used by the reference's
older tests and helpers; the Solver works on particle_dtype arrays)."""
import numpy as np


class Particle:
    def __init__(self, label: str, x: float, y: float, mass: float, rho: float = 1000):
        self.label = label
        self.r = np.array([x, y])
        self.m = mass
        self.rho = rho
        self.v = np.zeros(2)        # velocity
        self.vx = np.zeros(2)       # XSPH-corrected velocity
        self.a = np.zeros(2)        # acceleration
        self.p = 0.0                # pressure
        self.drho = 0.0             # density change

    def __eq__(self, other):
        return self.r[0] == other.r[0] and self.r[1] == other.r[1]
