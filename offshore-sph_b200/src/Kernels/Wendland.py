"""Wendland kernel as shipped, alpha = 9/(4 pi h^2), support q < 2 (reference src/Kernels/Wendland.py:9-64).

Inside Solver.run() the kernel is evaluated inline by the fused pair kernel (csrc/pair.cu); the
array methods below serve callers that use the kernel object on its own (e.g. the IceBreak pressure
probe) and run on the device through osph_leaf_kernel.
"""
import numpy as np

from src.Kernels.Kernel import Kernel
from osph_b200 import capi


class Wendland(Kernel):
    osph_name = 'wendland'

    @staticmethod
    def evaluate(r: np.array, h: np.array):
        return capi.leaf_kernel('wendland', 0, None, r, h)

    @staticmethod
    def gradient(x: np.array, r: np.array, h: np.array):
        return capi.leaf_kernel('wendland', 1, x, r, h)
