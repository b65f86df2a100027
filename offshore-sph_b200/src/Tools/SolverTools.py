"""Helpers with the names of reference src/Tools/SolverTools.py: findActive (:46-72), computeH (:106-118),
`_loop` (:120-174) as a stand-alone call of the fused device path, and `_assignProps` / `computeProps` (:8-44, :74-104),
which build the per-neighbour table of ONE particle for callers that use the equations one by one (the fused pair
kernel never materialises it)."""
from typing import Tuple

import numpy as np

from osph_b200 import capi


def findActive(J: int, pA: np.array) -> Tuple[int, np.array]:
    a_i = ~np.asarray(pA['deleted'], dtype=bool)
    return int(np.sum(a_i)), a_i


def computeH(sigma: float, J: int, m: np.array, rho: np.array):
    return capi.leaf_compute_h(sigma, m, rho)


def _assignProps(i: int, particleArray: np.array, near_arr: np.array, h_i: np.array, q_i: np.array, dist: np.array):
    """computed_dtype table of particle i over its neighbours near_arr: label / p / m / rho copied from the neighbours' rows,
    h / q / r from the neighbour query, and the differences i - j of x, y, vx, vy (device: osph_leaf_differences);
    c, w, dw_x, dw_y stay 0 as in the reference (:83-101)."""
    from src.Common import computed_dtype
    near = np.asarray(near_arr).astype(np.int64)
    comp = np.zeros(len(near), dtype=computed_dtype)
    if len(near) == 0:
        return comp
    rows = particleArray[near]
    for f in ('label', 'p', 'm', 'rho'):
        comp[f] = rows[f]
    comp['h'] = h_i; comp['q'] = q_i; comp['r'] = dist
    me = particleArray[i]
    d = capi.leaf_differences([me['x'], me['y'], me['vx'], me['vy']],
                              np.stack([rows['x'], rows['y'], rows['vx'], rows['vy']]))
    comp['x'], comp['y'], comp['vx'], comp['vy'] = d
    return comp


def computeProps(i: int, pA: np.array, near_arr, h_i, q_i, dist, evFunc, gradFunc):
    """_assignProps plus the kernel columns w, dw_x, dw_y (reference :8-44); evFunc / gradFunc are the evaluate / gradient
    of a src.Kernels class (device leaves)."""
    comp = _assignProps(i, pA, near_arr, h_i, q_i, dist)
    if len(comp):
        comp['w'] = evFunc(comp['r'], comp['h'])
        comp['dw_x'] = gradFunc(comp['x'], comp['r'], comp['h'])
        comp['dw_y'] = gradFunc(comp['y'], comp['r'], comp['h'])
    return comp


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
