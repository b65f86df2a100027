"""Monaghan-1992 predictor-corrector; constructor as reference src/Integrators/PEC.py:13-28.
The update formulas run in csrc/step.cu (k_prepare / k_correct), fused with the grid-bounds and
time-step reductions."""
from src.Integrators.Integrator import Integrator


class PEC(Integrator):
    osph_name = 'pec'

    def __init__(self, useXSPH: bool = True, strict: bool = True):
        self.useXSPH = bool(useXSPH)
        self.strict = bool(strict)
