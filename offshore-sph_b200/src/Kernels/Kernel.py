"""Base of the smoothing-kernel plugins (CubicSpline, Wendland, Gaussian).

The Solver only needs two things from a kernel object: which device kernel it selects (`osph_name`, one of the
OSPH_KERNEL_* names of include/osph.h) and, for code that uses the object on its own, the two array functions of
the reference interface (src/Kernels/Kernel.py:5-12): evaluate(r, h) -> W and gradient(x, r, h) -> dW/dx component.
Subclasses provide both as static methods that forward to the device leaf `osph_leaf_kernel`.
"""


class Kernel:
    osph_name = None                     # 'cubic' | 'wendland' | 'gaussian'

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        missing = [m for m in ("evaluate", "gradient") if m not in cls.__dict__]
        if missing:
            raise TypeError("%s must define %s" % (cls.__name__, " and ".join(missing)))

    @staticmethod
    def evaluate(r, h):
        raise NotImplementedError("Kernel.evaluate: use CubicSpline, Wendland or Gaussian")

    @staticmethod
    def gradient(x, r, h):
        raise NotImplementedError("Kernel.gradient: use CubicSpline, Wendland or Gaussian")
