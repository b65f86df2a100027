"""Newmark-beta integrator for Coupled (structural) particles; interface as reference
src/Integrators/NewmarkBeta.py:13-96.

Host-side by design: it advances the O(10^2..10^3) nodes of a user-supplied structural model inside the
coupling callback path and is outside the device hot path.
"""
import numpy as np

from src.Common import ParticleType


class NewmarkBeta:
    osph_name = None

    def __init__(self, beta: float, gamma: float, M: np.array, K: np.array, C: np.array):
        assert beta * 2 >= gamma
        self.beta = beta
        self.gamma = gamma
        self.M = np.asarray(M, dtype=np.float64)
        self.K = np.asarray(K, dtype=np.float64)
        self.C = np.asarray(C, dtype=np.float64)
        self._lhs_key = None
        self._lhs_inv = None

    def isMultiStage(self) -> bool:
        return False

    def _check(self, pA):
        if np.any(pA['label'] != ParticleType.Coupled):
            print('\nNewmark-Beta integration method is only suited for solid coupling particles.')

    def predict(self, dt: float, pA: np.array, damping: float = 0.0):
        self._check(pA)
        a, b = (0.5 - self.beta) * dt ** 2, (1 - self.gamma) * dt
        pA['x'] = pA['x'] + pA['vx'] * dt + pA['ax'] * a
        pA['y'] = pA['y'] + pA['vy'] * dt + pA['ay'] * a
        pA['vx'] = pA['vx'] + pA['ax'] * b
        pA['vy'] = pA['vy'] + pA['ay'] * b
        return pA

    def correct(self, dt: float, pA: np.array, damping: float = 0.0):
        self._check(pA)
        a, b = self.beta * dt ** 2, self.gamma * dt
        pA['x'] = pA['x'] + pA['ax'] * a
        pA['y'] = pA['y'] + pA['ay'] * a
        pA['vx'] = pA['vx'] + pA['ax'] * b
        pA['vy'] = pA['vy'] + pA['ay'] * b
        return pA

    def acceleration(self, dt: float, F: np.array, r: np.array, v: np.array):
        """a = (M + gamma dt C + beta dt^2 K)^-1 (F - K r - C v); the factorisation is reused while dt is unchanged."""
        rhs = F - self.K @ r - self.C @ v
        if self._lhs_key != dt:
            self._lhs_inv = np.linalg.inv(self.M + self.C * self.gamma * dt + self.K * self.beta * dt ** 2)
            self._lhs_key = dt
        return self._lhs_inv @ rhs
