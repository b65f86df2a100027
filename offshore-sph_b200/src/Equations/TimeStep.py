"""Time-step criteria, interface of reference src/Equations/TimeStep.py:8-91.
Inside Solver.run() the reduction is fused into the corrector kernel; this class serves stand-alone calls."""
import numpy as np


def _scratch_context(pA, integrator='pec', cfl=(0.25, 0.25)):
    from osph_b200 import capi
    from src.Common import particle_dtype
    arr = np.ascontiguousarray(pA).astype(particle_dtype, copy=True)
    arr['deleted'] = False
    consts = dict(height=1.0, r0=1.0, rho0=1000.0, Pb=0.0, gamma=7.0, co=1.0, B=1.0, alpha=0.0, beta=0.0,
                  epsilon=0.5, D=0.0, p1=4.0, p2=2.0, useXSPH=False)
    cfg = capi.make_config(consts, 'cubic', integrator, capi.FP64, None, keep_h=True, device=capi.default_device())
    cfg.cfl_courant, cfg.cfl_force = cfl
    ctx = capi.Context(cfg)
    ctx.upload(arr)
    return ctx


class TimeStep:
    def compute(self, J: int, pA: np.array, gamma_c: float = 0.25, gamma_f: float = 0.25):
        """(min, courant, force) over the fluid rows of pA; J is ignored like in the reference."""
        ctx = _scratch_context(pA, cfl=(gamma_c, gamma_f))
        try:
            return ctx.timestep()
        finally:
            ctx.close()

    def courant(self, cfl, h_min, c_max) -> float:
        return cfl * h_min / c_max

    def force(self, cfl, min_h, max_a) -> float:
        return 1e10 if max_a < 1e-12 else cfl * float(np.sqrt(min_h / max_a))
