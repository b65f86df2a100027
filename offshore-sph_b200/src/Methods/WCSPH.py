"""WCSPH method plugin: parameter holder + dispatch tag for the device path.

Constructor and attributes mirror reference src/Methods/WCSPH.py:35-79.  The per-pair arithmetic
(compute_density_change / compute_acceleration / compute_velocity in the reference) lives in the fused
pair kernel; called stand-alone on one particle and its computed-neighbour table, those three run on the device
through osph_leaf_equations, as do the three whole-array methods the Solver calls during setup.
"""
import math

import numpy as np

from src.Methods.Method import Method
from osph_b200 import capi


class WCSPH(Method):
    def __init__(self, height: float, r0: float, rho0: float, useXSPH: bool, Pb: float = 0.0,
                 useSummationDensity: bool = False):
        # useSummationDensity defaults to False so that examples/Containment.py, which omits it, runs.
        self.height = float(height)
        self.rho0 = float(rho0)
        self.useXSPH = bool(useXSPH)
        self.useSummationDensity = bool(useSummationDensity)
        self.epsilon = 0.5
        self.gamma = 7.0
        self.co = 10.0 * math.sqrt(2 * 9.81 * self.height)
        self.B = self.co * self.co * self.rho0 / self.gamma
        self.Pb = float(Pb)
        self.alpha = 0.01
        self.beta = 0.0
        self.r0 = float(r0)
        self.D = 5 * 9.81 * self.height
        self.p1 = 4
        self.p2 = 2

    def constants(self):
        return dict(height=self.height, r0=self.r0, rho0=self.rho0, Pb=self.Pb, gamma=self.gamma, co=self.co,
                    B=self.B, alpha=self.alpha, beta=self.beta, epsilon=self.epsilon, D=self.D,
                    p1=float(self.p1), p2=float(self.p2), useXSPH=self.useXSPH,
                    useSummationDensity=self.useSummationDensity)

    def initialize(self, pA: np.array):
        pA['rho'] = capi.leaf_tait_height(pA['y'], self.rho0, self.height, self.B, self.gamma)
        return pA

    def compute_speed_of_sound(self, pA: np.array) -> np.array:
        return np.full(len(pA), self.co, dtype=np.float64)

    def compute_pressure(self, pA: np.array) -> np.array:
        return capi.leaf_tait_pressure(pA['rho'], pA['label'], self.gamma, self.B, self.rho0, self.Pb)

    # ---- per-particle forms (reference WCSPH.py:151-203): one particle record + its computed-neighbour table ----
    def compute_acceleration(self, p: np.array, comp: np.array):
        a = capi.leaf_equations(p, comp, alpha=self.alpha, beta=self.beta)['a']
        return [a[0], a[1] - 9.81]

    def compute_velocity(self, p: np.array, comp: np.array):
        if self.useXSPH:
            xs = capi.leaf_equations(p, comp, epsilon=self.epsilon)['xsph']
            return [p['vx'], p['vy'], p['vx'] + xs[0], p['vy'] + xs[1]]
        return [p['vx'], p['vy'], 0.0, 0.0]

    def compute_density_change(self, p: np.array, comp: np.array):
        if self.useSummationDensity:
            return 0.0
        return capi.leaf_equations(p, comp)['drho']
