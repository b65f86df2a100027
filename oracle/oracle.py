"""ctypes front end of the CPU oracle (oracle/wcsph_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.

The oracle works on SoA views of the ACTIVE rows of a ``particle_dtype`` array
(reference src/Common.py:26-57: packed, itemsize 154).  `Particles.from_aos`
copies the columns out, `to_aos` writes them back.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

FIELDS = ['m', 'rho', 'p', 'c', 'drho', 'h', 'x', 'y', 'vx', 'vy', 'ax', 'ay',
          'xsphx', 'xsphy', 'x0', 'y0', 'vx0', 'vy0', 'rho0']

# Host interchange record of the reference (src/Common.py:26-57): bool, int8, 19 packed doubles.
particle_dtype = np.dtype({'names': ['deleted', 'label'] + FIELDS,
                           'formats': [np.bool_, np.int8] + [np.float64] * len(FIELDS)})
assert particle_dtype.itemsize == 154

KERNELS = {'cubic': 0, 'wendland': 1, 'gaussian': 2}
INTEGRATORS = {'pec': 0, 'euler': 1, 'verlet': 2}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


class _CParticles(C.Structure):
    _fields_ = [('n', C.c_int64), ('label', C.POINTER(C.c_int8))] + [(f, _dp) for f in FIELDS]


class WCSPHParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ('height', 'r0', 'rho0', 'Pb', 'gamma', 'co', 'B', 'alpha', 'beta',
                 'epsilon', 'D', 'p1', 'p2')] + [('useXSPH', C.c_int), ('useSummationDensity', C.c_int)]


class _CGrid(C.Structure):
    _fields_ = [(k, C.c_double) for k in ('xmin', 'xmax', 'ymin', 'ymax', 'cell_size', 'scale')] + \
               [(k, C.c_int64) for k in ('ncx', 'ncy', 'n_cells', 'n')] + \
               [('heads', _ip), ('nexts', _ip)]


def build(force=False):
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "wcsph_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_tait_co.restype = C.c_double; L.oracle_tait_co.argtypes = [C.c_double]
        L.oracle_tait_B.restype = C.c_double; L.oracle_tait_B.argtypes = [C.c_double] * 3
        L.oracle_tait_p.restype = C.c_double
        L.oracle_tait_p.argtypes = [C.c_double] * 4 + [C.c_int]
        L.oracle_tait_height.restype = C.c_double; L.oracle_tait_height.argtypes = [C.c_double] * 5
        L.oracle_wcsph_init.argtypes = [C.POINTER(WCSPHParams), C.c_double, C.c_double, C.c_double,
                                        C.c_int, C.c_double, C.c_int]
        L.oracle_wcsph_initialize.argtypes = [C.POINTER(WCSPHParams), C.c_int64, _dp, _dp]
        L.oracle_compute_h.argtypes = [C.c_double, C.c_int64, _dp, _dp, _dp]
        L.oracle_kernel_evaluate.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp]
        L.oracle_kernel_gradient.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp, _dp]
        L.oracle_nn_update.argtypes = [C.POINTER(_CGrid), C.c_double, C.c_int64, _dp, _dp, _dp]
        L.oracle_grid_free.argtypes = [C.POINTER(_CGrid)]
        L.oracle_cell_ids.argtypes = [C.POINTER(_CGrid), C.c_int64, _dp, _dp, _ip]
        L.oracle_near_pos.restype = C.c_int64
        L.oracle_near_pos.argtypes = [C.POINTER(_CGrid), C.c_double, C.c_double, C.c_double,
                                      _dp, _dp, _dp, C.c_int64, _ip, _dp, _dp, _dp]
        L.oracle_neighbours_csr.restype = C.c_int64
        L.oracle_neighbours_csr.argtypes = [C.POINTER(_CGrid), C.c_int64, C.POINTER(C.c_int8),
                                            _dp, _dp, _dp, _ip, C.c_int64, _ip]
        L.oracle_loop.restype = C.c_int64
        L.oracle_loop.argtypes = [C.POINTER(_CParticles), C.POINTER(WCSPHParams), C.POINTER(_CGrid),
                                  C.c_int, C.c_int64, C.c_int64]
        L.oracle_loop_mt.restype = C.c_int64
        L.oracle_loop_mt.argtypes = [C.POINTER(_CParticles), C.POINTER(WCSPHParams), C.POINTER(_CGrid),
                                     C.c_int, C.c_int64, C.c_int64, C.c_int]
        u8 = C.POINTER(C.c_uint8)
        L.oracle_pec_predict.argtypes = [C.POINTER(_CParticles), u8, C.c_double, C.c_double, C.c_int, C.c_int]
        L.oracle_pec_correct.argtypes = [C.POINTER(_CParticles), u8, C.c_double, C.c_double, C.c_int, C.c_int]
        L.oracle_euler_correct.argtypes = [C.POINTER(_CParticles), u8, C.c_double]
        L.oracle_verlet_predict.argtypes = [C.POINTER(_CParticles), u8, C.c_double]
        L.oracle_verlet_correct.argtypes = [C.POINTER(_CParticles), u8, C.c_double, C.c_int]
        L.oracle_timestep.restype = C.c_int
        L.oracle_timestep.argtypes = [C.POINTER(_CParticles), u8, C.c_double, C.c_double, _dp]
        L.oracle_kinetic_energy.restype = C.c_double
        L.oracle_kinetic_energy.argtypes = [C.POINTER(_CParticles), u8]
        L.oracle_step.restype = C.c_int
        L.oracle_step.argtypes = [C.POINTER(_CParticles), C.POINTER(WCSPHParams), C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_double, C.c_double, C.c_double, u8, C.c_int64, _dp, _ip]
        i8 = C.POINTER(C.c_int8)
        L.oracle_eq_continuity.restype = C.c_double
        L.oracle_eq_continuity.argtypes = [C.c_int64, i8] + [_dp] * 5
        L.oracle_eq_momentum.argtypes = [C.c_double] * 6 + [C.c_int64, i8] + [_dp] * 13
        L.oracle_eq_xsph.argtypes = [C.c_double, C.c_double, C.c_int64] + [_dp] * 6
        L.oracle_eq_boundary_force.argtypes = [C.c_double] * 4 + [C.c_int64, i8] + [_dp] * 4
        L.oracle_eq_courant.restype = C.c_double
        L.oracle_eq_courant.argtypes = [C.c_double, C.c_int64, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _u8(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


def wcsph(height, r0, rho0, useXSPH, Pb=0.0, useSummationDensity=False):
    """WCSPH constants, reference src/Methods/WCSPH.py:35-79."""
    w = WCSPHParams()
    lib().oracle_wcsph_init(C.byref(w), height, r0, rho0, int(bool(useXSPH)), Pb, int(bool(useSummationDensity)))
    return w


class Particles:
    """SoA copy of the active rows of a particle_dtype array."""

    def __init__(self, n):
        self.n = n
        self.label = np.zeros(n, dtype=np.int8)
        for f in FIELDS:
            setattr(self, f, np.zeros(n, dtype=np.float64))
        self._c = None

    @classmethod
    def from_aos(cls, pA):
        P = cls(len(pA))
        P.label[:] = pA['label']
        for f in FIELDS:
            getattr(P, f)[:] = pA[f]
        return P

    def to_aos(self, pA=None):
        if pA is None:
            pA = np.zeros(self.n, dtype=particle_dtype)
        pA['label'] = self.label
        for f in FIELDS:
            pA[f] = getattr(self, f)
        return pA

    def copy(self):
        Q = Particles(self.n)
        Q.label[:] = self.label
        for f in FIELDS:
            getattr(Q, f)[:] = getattr(self, f)
        return Q

    @property
    def fluid(self):
        return (self.label == 0).astype(np.uint8)

    def cref(self):
        s = _CParticles()
        s.n = self.n
        s.label = self.label.ctypes.data_as(C.POINTER(C.c_int8))
        for f in FIELDS:
            setattr(s, f, _d(getattr(self, f)))
        self._c = s
        return C.byref(s)


class Grid:
    """The reference's NNLinkedList state after update() (src/Tools/NNLinkedList.py:37-39)."""

    def __init__(self, P, scale=2.0):
        self.P = P
        self.g = _CGrid()
        self.rc = lib().oracle_nn_update(C.byref(self.g), scale, P.n, _d(P.x), _d(P.y), _d(P.h))

    def __del__(self):
        try:
            lib().oracle_grid_free(C.byref(self.g))
        except Exception:
            pass

    @property
    def params(self):
        g = self.g
        return dict(xmin=g.xmin, xmax=g.xmax, ymin=g.ymin, ymax=g.ymax, cell_size=g.cell_size,
                    ncx=g.ncx, ncy=g.ncy, n_cells=g.n_cells)

    def cell_ids(self):
        out = np.zeros(self.P.n, dtype=np.int64)
        lib().oracle_cell_ids(C.byref(self.g), self.P.n, _d(self.P.x), _d(self.P.y),
                              out.ctypes.data_as(_ip))
        return out

    def near_pos(self, x, y, h):
        P = self.P
        cap = 64
        while True:
            idx = np.zeros(cap, dtype=np.int64); r = np.zeros(cap); q = np.zeros(cap); hh = np.zeros(cap)
            cnt = lib().oracle_near_pos(C.byref(self.g), x, y, h, _d(P.x), _d(P.y), _d(P.h), cap,
                                        idx.ctypes.data_as(_ip), _d(r), _d(q), _d(hh))
            if cnt <= cap:
                return hh[:cnt], q[:cnt], r[:cnt], idx[:cnt]
            cap = int(cnt)

    def near(self, i):
        P = self.P
        return self.near_pos(float(P.x[i]), float(P.y[i]), float(P.h[i]))

    def neighbours_csr(self):
        P = self.P
        off = np.zeros(P.n + 1, dtype=np.int64)
        lab = P.label.ctypes.data_as(C.POINTER(C.c_int8))
        total = lib().oracle_neighbours_csr(C.byref(self.g), P.n, lab, _d(P.x), _d(P.y), _d(P.h),
                                            off.ctypes.data_as(_ip), 0, None)
        idx = np.zeros(max(int(total), 1), dtype=np.int64)
        lib().oracle_neighbours_csr(C.byref(self.g), P.n, lab, _d(P.x), _d(P.y), _d(P.h),
                                    off.ctypes.data_as(_ip), total, idx.ctypes.data_as(_ip))
        return off, idx[:total]


def loop(P, w, grid, kernel='cubic', stride=1, phase=0, threads=1):
    """Reference `_loop` (src/Tools/SolverTools.py:120-174), in place on P. Returns #pairs.
    threads > 1: the pair loop on that many host threads (same arithmetic per particle, bit-identical results)."""
    if threads > 1:
        return lib().oracle_loop_mt(P.cref(), C.byref(w), C.byref(grid.g), KERNELS[kernel], stride, phase, int(threads))
    return lib().oracle_loop(P.cref(), C.byref(w), C.byref(grid.g), KERNELS[kernel], stride, phase)


def pec_predict(P, mask, dt, damping, useXSPH=True, strict=False):
    lib().oracle_pec_predict(P.cref(), _u8(mask), dt, damping, int(useXSPH), int(strict))


def pec_correct(P, mask, dt, damping, useXSPH=True, strict=False):
    lib().oracle_pec_correct(P.cref(), _u8(mask), dt, damping, int(useXSPH), int(strict))


def euler_correct(P, mask, dt):
    lib().oracle_euler_correct(P.cref(), _u8(mask), dt)


def verlet_predict(P, mask, dt):
    lib().oracle_verlet_predict(P.cref(), _u8(mask), dt)


def verlet_correct(P, mask, dt, useXSPH=True):
    lib().oracle_verlet_correct(P.cref(), _u8(mask), dt, int(useXSPH))


def timestep(P, mask, gamma_c=0.25, gamma_f=0.25):
    out = np.zeros(3)
    rc = lib().oracle_timestep(P.cref(), _u8(mask), gamma_c, gamma_f, _d(out))
    if rc != 0:
        raise ValueError("no fluid particles")
    return tuple(out)


def kinetic_energy(P, mask=None):
    return lib().oracle_kinetic_energy(P.cref(), _u8(mask))


def kernel_evaluate(kernel, r, h):
    r = np.ascontiguousarray(r, dtype=np.float64); h = np.ascontiguousarray(h, dtype=np.float64)
    out = np.zeros_like(r)
    lib().oracle_kernel_evaluate(KERNELS[kernel], len(r), _d(r), _d(h), _d(out))
    return out


def kernel_gradient(kernel, x, r, h):
    x = np.ascontiguousarray(x, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64); h = np.ascontiguousarray(h, dtype=np.float64)
    out = np.zeros_like(r)
    lib().oracle_kernel_gradient(KERNELS[kernel], len(r), _d(x), _d(r), _d(h), _d(out))
    return out


# ---- the equations as leaves on a computed-neighbour table (numpy structured array with the reference's
# computed_dtype field names: label p rho x y vx vy r h c m w dw_x dw_y) ----
def _col(comp, f):
    return np.ascontiguousarray(comp[f], dtype=np.float64)


def _lab(comp):
    return np.ascontiguousarray(comp['label'], dtype=np.int8)


def eq_continuity(comp):
    lab = _lab(comp)
    cols = [_col(comp, f) for f in ('m', 'vx', 'vy', 'dw_x', 'dw_y')]
    return float(lib().oracle_eq_continuity(len(comp), lab.ctypes.data_as(C.POINTER(C.c_int8)), *[_d(a) for a in cols]))


def eq_momentum(alpha, beta, p, comp):
    """p: mapping with the self particle's p, rho, h, c (Momentum.py:6-57)."""
    lab = _lab(comp)
    cols = [_col(comp, f) for f in ('p', 'rho', 'x', 'y', 'vx', 'vy', 'r', 'h', 'c', 'm', 'dw_x', 'dw_y')]
    out = np.zeros(2)
    lib().oracle_eq_momentum(alpha, beta, float(p['p']), float(p['rho']), float(p['h']), float(p['c']), len(comp),
                             lab.ctypes.data_as(C.POINTER(C.c_int8)), *[_d(a) for a in cols], _d(out))
    return out


def eq_xsph(epsilon, p, comp):
    cols = [_col(comp, f) for f in ('rho', 'm', 'w', 'vx', 'vy')]
    out = np.zeros(2)
    lib().oracle_eq_xsph(epsilon, float(p['rho']), len(comp), *[_d(a) for a in cols], _d(out))
    return out


def eq_boundary_force(r0, D, p1, p2, comp):
    lab = _lab(comp)
    cols = [_col(comp, f) for f in ('r', 'x', 'y')]
    out = np.zeros(2)
    lib().oracle_eq_boundary_force(r0, D, p1, p2, len(comp), lab.ctypes.data_as(C.POINTER(C.c_int8)),
                                   *[_d(a) for a in cols], _d(out))
    return out


def eq_courant(alpha, h, c):
    h = np.ascontiguousarray(h, dtype=np.float64); c = np.ascontiguousarray(c, dtype=np.float64)
    return float(lib().oracle_eq_courant(alpha, len(h), _d(h), _d(c)))


def compute_h(sigma, m, rho):
    m = np.ascontiguousarray(m, dtype=np.float64); rho = np.ascontiguousarray(rho, dtype=np.float64)
    out = np.zeros_like(m)
    lib().oracle_compute_h(sigma, len(m), _d(m), _d(rho), _d(out))
    return out


def initialize_density(w, y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros_like(y)
    lib().oracle_wcsph_initialize(C.byref(w), len(y), _d(y), _d(out))
    return out


def step(P, w, kernel='cubic', integrator='pec', integ_xsph=True, strict=False, damping=0.05,
         fixed_h=None, fixed_dt=None, stride=1):
    """One whole step in the order of src/Solver.py:366-399, in place. Returns ((dt,dt_c,dt_f), pairs)."""
    dt3 = np.zeros(3)
    pairs = C.c_int64(0)
    fluid = P.fluid
    rc = lib().oracle_step(P.cref(), C.byref(w), KERNELS[kernel], INTEGRATORS[integrator], int(integ_xsph),
                           int(strict), damping, -1.0 if fixed_h is None else fixed_h,
                           -1.0 if fixed_dt is None else fixed_dt, _u8(fluid), stride, _d(dt3),
                           C.byref(pairs))
    if rc != 0:
        raise ValueError("oracle_step failed (no fluid particles)")
    return tuple(dt3), pairs.value
