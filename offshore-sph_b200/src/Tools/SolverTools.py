"""Helpers with the names of reference src/Tools/SolverTools.py: findActive (:46-72), computeH (:106-118)
and `_loop` (:120-174), the latter as a stand-alone call of the fused device path."""
from typing import Tuple

import numpy as np

from osph_b200 import capi


def findActive(J: int, pA: np.array) -> Tuple[int, np.array]:
    a_i = ~np.asarray(pA['deleted'], dtype=bool)
    return int(np.sum(a_i)), a_i


def computeH(sigma: float, J: int, m: np.array, rho: np.array):
    return capi.leaf_compute_h(sigma, m, rho)


def _kernel_name(evFunc):
    owner = getattr(evFunc, '__qualname__', '').split('.')[0]
    try:
        return {'CubicSpline': 'cubic', 'Wendland': 'wendland', 'Gaussian': 'gaussian'}[owner]
    except KeyError:
        raise TypeError('kernel functions must come from src.Kernels.{CubicSpline,Wendland,Gaussian}')


def _loop(pA, evFunc, gradFunc, methodClass, nn=None):
    """One force evaluation of the active array pA: EOS + fused pair interactions, h left as passed in."""
    from src.Common import particle_dtype
    arr = np.ascontiguousarray(pA).astype(particle_dtype, copy=True)
    arr['deleted'] = False
    cfg = capi.make_config(methodClass.constants(), _kernel_name(evFunc), 'pec', capi.precision_from_env(), None,
                           keep_h=True, device=capi.default_device())
    if nn is not None:
        cfg.nn_scale = nn.scale
    with capi.Context(cfg) as ctx:
        ctx.upload(arr)
        ctx.compute()
        ctx.download(arr)
    arr['deleted'] = pA['deleted']
    pA[...] = arr
    return pA
