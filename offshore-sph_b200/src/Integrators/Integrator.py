"""Integrator base class (reference src/Integrators/Integrator.py:4-36)."""
import abc

import numpy as np


class Integrator(abc.ABC):
    #: name understood by the C ABI (OSPH_INTEGRATOR_*); None = host-side integrator
    osph_name = None
    useXSPH = False
    strict = False

    def isMultiStage(self) -> bool:
        return False

    def _standalone(self, phase, dt, pA, damping):
        """predict/correct on a host array outside Solver.run(): upload, run the device kernel, download.

        Like the reference, the rows passed in are all integrated (the Solver only ever passes fluid rows);
        the device kernels act on fluid-labelled rows, so other labels are relabelled for the call.
        """
        from osph_b200 import capi
        from src.Common import particle_dtype
        arr = np.ascontiguousarray(np.atleast_1d(pA)).astype(particle_dtype, copy=True)
        labels = arr['label'].copy()
        deleted = arr['deleted'].copy()
        arr['label'] = 0
        arr['deleted'] = False
        consts = dict(height=1.0, r0=1.0, rho0=1000.0, Pb=0.0, gamma=7.0, co=1.0, B=1.0, alpha=0.0, beta=0.0,
                      epsilon=0.5, D=0.0, p1=4.0, p2=2.0, useXSPH=self.useXSPH)
        cfg = capi.make_config(consts, 'cubic', self.osph_name, capi.FP64, None, integrator_xsph=self.useXSPH,
                               strict=self.strict, keep_h=True, device=capi.default_device())
        with capi.Context(cfg) as ctx:
            ctx.upload(arr)
            if phase == 'predict':
                ctx.predict(dt, damping)
            else:
                ctx.correct(dt, damping)
            ctx.download(arr)
        arr['label'] = labels
        arr['deleted'] = deleted
        if isinstance(pA, np.ndarray) and pA.shape == arr.shape:
            pA[...] = arr
            return pA
        return arr

    def predict(self, dt: float, pA: np.array, damping: float = 0.0):
        return self._standalone('predict', dt, pA, damping)

    def correct(self, dt: float, pA: np.array, damping: float = 0.0):
        return self._standalone('correct', dt, pA, damping)
