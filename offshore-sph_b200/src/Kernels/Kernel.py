"""Kernel base class; same two-method interface as reference src/Kernels/Kernel.py:5-12."""
import abc

import numpy as np


class Kernel(metaclass=abc.ABCMeta):
    #: name understood by the C ABI (OSPH_KERNEL_*)
    osph_name = None

    @staticmethod
    @abc.abstractmethod
    def evaluate(r: np.array, h: np.array) -> np.array:
        pass

    @staticmethod
    @abc.abstractmethod
    def gradient(x: np.array, r: np.array, h: np.array) -> np.array:
        pass
