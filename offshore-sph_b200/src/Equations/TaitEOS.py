"""Tait equation of state, call signatures of reference src/Equations/TaitEOS.py:6-65.
Array functions run on the device (osph_leaf_tait_pressure / osph_leaf_tait_height); inside the step the
EOS is fused into the gather kernel (csrc/step.cu: k_gather)."""
from math import sqrt

import numpy as np

from osph_b200 import capi


def TaitEOS(gamma, B, rho0, rho, label):
    """B ((rho/rho0)^gamma - 1) for fluid rows, 0 otherwise."""
    scalar = np.ndim(rho) == 0
    out = capi.leaf_tait_pressure(np.atleast_1d(rho), np.atleast_1d(label), gamma, B, rho0, 0.0)
    return float(out[0]) if scalar else out


def TaitEOS_B(co, rho0, gamma):
    return co * co * rho0 / gamma


def TaitEOS_co(H):
    return 10.0 * sqrt(2 * 9.81 * H)


def TaitEOS_height(rho0, H, B, gamma, y):
    scalar = np.ndim(y) == 0
    out = capi.leaf_tait_height(np.atleast_1d(y), rho0, H, B, gamma)
    return float(out[0]) if scalar else out
