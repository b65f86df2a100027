"""2-D Gaussian cut at q = 3, alpha = 1/(pi h^2) (reference src/Kernels/Gaussian.py:16-59).

Inside Solver.run() the kernel is evaluated inline by the fused pair kernel (csrc/pair.cu); the
array methods below serve callers that use the kernel object on its own (e.g. the IceBreak pressure
probe) and run on the device through osph_leaf_kernel.
"""
import numpy as np

from src.Kernels.Kernel import Kernel
from osph_b200 import capi


class Gaussian(Kernel):
    osph_name = 'gaussian'

    @staticmethod
    def evaluate(r: np.array, h: np.array):
        return capi.leaf_kernel('gaussian', 0, None, r, h)

    @staticmethod
    def gradient(x: np.array, r: np.array, h: np.array):
        return capi.leaf_kernel('gaussian', 1, x, r, h)

    @staticmethod
    def derivative(r: np.array, h: np.array):
        """dW/dq (reference Gaussian.py:27-36) = gradient along the separation times h."""
        r = np.asarray(r, dtype=np.float64)
        return capi.leaf_kernel('gaussian', 1, r, r, h) * np.asarray(h, dtype=np.float64)
