"""ctypes binding of libosph_b200.so (include/osph.h).

This is the thin C-ABI layer named by the north star: host Python hands the
reference's packed ``particle_dtype`` array to hand-written sm_100a kernels and
gets it back.  There is no CPU fallback anywhere in this module: if the shared
library is missing or no CUDA device is present, construction raises.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_PATH = os.environ.get("OSPH_LIB") or os.path.join(PKG_ROOT, "lib", "libosph_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_ROOT), "include", "osph.h")

FP64, FP32 = 0, 1
KERNELS = {'cubic': 0, 'wendland': 1, 'gaussian': 2}
INTEGRATORS = {'pec': 0, 'euler': 1, 'verlet': 2}
FIELDS = ['m', 'rho', 'p', 'c', 'drho', 'h', 'x', 'y', 'vx', 'vy', 'ax', 'ay',
          'xsphx', 'xsphy', 'x0', 'y0', 'vx0', 'vy0', 'rho0']
FIELD_ID = {f: i for i, f in enumerate(FIELDS)}

S_NONFINITE, S_SMALL_DT, S_UNBINNED, S_GRID_COARSE, S_DENSE_CELL = 1, 2, 4, 8, 16
S_SKIN_EXHAUSTED, S_H_NOT_UNIFORM = 32, 64


class OsphError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libosph_b200 error %d: %s" % (code, text))
        self.code = code


class Config(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ('struct_size', 'device', 'precision', 'kernel', 'integrator', 'method_xsph',
                 'integrator_xsph', 'strict', 'summation_density', 'dynamic_h', 'reorder_every', 'reserved0')] + \
               [(k, C.c_double) for k in
                ('fixed_h', 'h_sigma', 'nn_scale', 'gamma', 'B', 'rho0', 'Pb', 'co', 'alpha', 'beta', 'epsilon',
                 'r0', 'D', 'p1', 'p2', 'gravity', 'cfl_courant', 'cfl_force', 'height')]


def build(force=False, verbose=False):
    """Compile libosph_b200.so for sm_100a with nvcc (csrc/Makefile). Works without a GPU."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))] + [HEADER]
    stale = force or not os.path.exists(LIB_PATH) or \
        os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    if stale:
        out = subprocess.run(["make", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout + out.stderr)
        if out.returncode != 0:
            raise RuntimeError("nvcc build of libosph_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (never builds implicitly on a box without the sources' toolchain)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("libosph_b200.so not found at %s: run __graft_entry__.build() (nvcc, sm_100a). "
                      "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    ctx = C.c_void_p
    i64, i32, dbl = C.c_int64, C.c_int32, C.c_double
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    sig = {
        'osph_version': (C.c_int, []),
        'osph_default_config': (C.c_int, [C.POINTER(Config), dbl, dbl, dbl]),
        'osph_create': (C.c_int, [C.POINTER(Config), C.POINTER(ctx)]),
        'osph_destroy': (C.c_int, [ctx]),
        'osph_last_error': (C.c_char_p, [ctx]),
        'osph_upload_aos': (C.c_int, [ctx, C.c_void_p, i64, i64]),
        'osph_download_aos': (C.c_int, [ctx, C.c_void_p, i64, i64]),
        'osph_import_device_aos': (C.c_int, [ctx, C.c_void_p, i64, i64]),
        'osph_export_device_aos': (C.c_int, [ctx, C.c_void_p, i64, i64]),
        'osph_download_fields': (C.c_int, [ctx, i32, C.POINTER(i32), C.POINTER(dp)]),
        'osph_upload_fields': (C.c_int, [ctx, i32, C.POINTER(i32), C.POINTER(dp)]),
        'osph_export_begin': (C.c_int, [ctx, i32, C.POINTER(i32), i32, ip]),
        'osph_export_end': (C.c_int, [ctx, i64, i32, C.POINTER(dp), i64]),
        'osph_download_rows': (C.c_int, [ctx, i64, ip, C.c_void_p, i64]),
        'osph_upload_rows': (C.c_int, [ctx, i64, ip, C.c_void_p, i64]),
        'osph_set_active': (C.c_int, [ctx, C.c_void_p, i64]),
        'osph_sort_stats': (C.c_int, [ctx, C.POINTER(C.c_int64)]),
        'osph_num_active': (i64, [ctx]),
        'osph_num_fluid': (i64, [ctx]),
        'osph_initialize': (C.c_int, [ctx]),
        'osph_timestep': (C.c_int, [ctx, dp]),
        'osph_predict': (C.c_int, [ctx, dbl, dbl]),
        'osph_build_neighbours': (C.c_int, [ctx]),
        'osph_compute': (C.c_int, [ctx]),
        'osph_correct': (C.c_int, [ctx, dbl, dbl]),
        'osph_step': (C.c_int, [ctx, i32, dbl, dbl]),
        'osph_get_dt_log': (C.c_int, [ctx, dp, i64, ip]),
        'osph_kinetic_energy': (C.c_int, [ctx, dp]),
        'osph_sync': (C.c_int, [ctx, C.POINTER(C.c_uint32)]),
        'osph_get_cells': (C.c_int, [ctx, dp, ip]),
        'osph_get_neighbours_csr': (C.c_int, [ctx, ip, ip, i64, ip]),
        'osph_near_pos': (C.c_int, [ctx, dbl, dbl, dbl, i64, ip, dp, dp, dp, ip]),
        'osph_probe_pressure': (C.c_int, [ctx, i64, dp, dp, dbl, dp, dp]),
        'osph_get_timers': (C.c_int, [ctx, dp]),
        'osph_launch_count': (i64, [ctx]),
        'osph_stream': (C.c_uint64, [ctx]),
        'osph_pair_kernel_time': (C.c_int, [ctx, dp, ip]),
        'osph_pair_kernel_info': (C.c_int, [ctx, C.POINTER(C.c_int64)]),
        'osph_reserve': (C.c_int, [ctx, i64]),
        'osph_set_row_ids': (C.c_int, [ctx, C.POINTER(i32), i64]),
        'osph_slab_configure': (C.c_int, [ctx, dbl, dbl, C.c_void_p, i64]),
        'osph_slab_step_plan': (C.c_int, [ctx, i32, i32]),
        'osph_slab_dt_local': (C.c_int, [ctx, C.c_void_p]),
        'osph_slab_step_begin': (C.c_int, [ctx, C.c_void_p, dbl, dbl]),
        'osph_slab_pack': (C.c_int, [ctx, dbl, C.c_void_p, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64, C.c_void_p]),
        'osph_slab_commit': (C.c_int, [ctx, i64, C.c_void_p, i64, i64, dp]),
        'osph_slab_step_end': (C.c_int, [ctx, dbl]),
        'osph_download_owned': (C.c_int, [ctx, C.c_void_p, i64, i64, C.POINTER(i32), ip]),
        'osph_slab_export': (C.c_int, [ctx, C.c_void_p, C.c_void_p, i32, C.POINTER(i32), C.POINTER(C.c_void_p)]),
        'osph_nccl_unique_id': (C.c_int, [C.c_char_p, C.c_char_p]),
        'osph_nccl_last_error': (C.c_char_p, []),
        'osph_slab_comm_create': (C.c_int, [ctx, C.c_char_p, C.c_char_p, C.c_int, C.c_int, dbl, dbl, dbl, dbl, i64, i64,
                                            C.POINTER(C.c_void_p)]),
        'osph_slab_comm_destroy': (C.c_int, [ctx, C.c_void_p]),
        'osph_slab_comm_attach': (C.c_int, [ctx, C.c_void_p]),
        'osph_slab_comm_set_bounds': (C.c_int, [ctx, C.c_void_p, dbl, dbl]),
        'osph_slab_run': (C.c_int, [ctx, C.c_void_p, i32, dbl, dbl]),
        'osph_slab_last_counts': (C.c_int, [C.c_void_p, ip]),
        'osph_slab_p2p_create': (C.c_int, [ctx, C.c_int, C.c_int, dbl, dbl, dbl, dbl, i64, i64, C.POINTER(C.c_void_p),
                                           C.c_char_p]),
        'osph_slab_p2p_connect': (C.c_int, [ctx, C.c_void_p, C.c_char_p]),
        'osph_slab_p2p_destroy': (C.c_int, [ctx, C.c_void_p]),
        'osph_slab_p2p_attach': (C.c_int, [ctx, C.c_void_p]),
        'osph_slab_p2p_set_bounds': (C.c_int, [ctx, C.c_void_p, dbl, dbl]),
        'osph_slab_p2p_run': (C.c_int, [ctx, C.c_void_p, i32, dbl, dbl]),
        'osph_slab_p2p_last_counts': (C.c_int, [C.c_void_p, ip]),
        'osph_slab_p2p_stats': (C.c_int, [C.c_void_p, ip]),
        'osph_leaf_kernel': (C.c_int, [C.c_int, C.c_int, C.c_int, i64, dp, dp, dp, dp]),
        'osph_leaf_tait_pressure': (C.c_int, [C.c_int, i64, dp, C.POINTER(C.c_int8), dbl, dbl, dbl, dbl, dp]),
        'osph_leaf_tait_height': (C.c_int, [C.c_int, i64, dp, dbl, dbl, dbl, dbl, dp]),
        'osph_leaf_compute_h': (C.c_int, [C.c_int, i64, dbl, dp, dp, dp]),
        'osph_leaf_equations': (C.c_int, [C.c_int, i64, C.POINTER(C.c_int8), dp] + [dbl] * 11 + [dp]),
        'osph_leaf_courant': (C.c_int, [C.c_int, dbl, i64, dp, dp, dp]),
        'osph_leaf_differences': (C.c_int, [C.c_int, i64, dp, dp, dp]),
        'osph_leaf_last_error': (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._osph_signatures = sig
    _lib = L
    return L


def default_config(height, r0, rho0=1000.0):
    cfg = Config()
    rc = lib().osph_default_config(C.byref(cfg), float(height), float(r0), float(rho0))
    if rc != 0:
        raise OsphError(rc, "osph_default_config")
    return cfg


H_FIXED, H_DYNAMIC, H_KEEP = 0, 1, 2


def make_config(consts, kernel='cubic', integrator='pec', precision=FP64, fixed_h=None, integrator_xsph=None,
                strict=False, device=0, reorder_every=0, keep_h=False):
    """Config from a dict of WCSPH constants (workloads.wcsph_constants / a src.Methods.WCSPH object's fields)."""
    cfg = default_config(consts['height'], consts['r0'], consts['rho0'])
    for k in ('gamma', 'B', 'Pb', 'co', 'alpha', 'beta', 'epsilon', 'D', 'p1', 'p2'):
        setattr(cfg, k, float(consts[k]))
    cfg.device = device
    cfg.precision = precision
    cfg.kernel = KERNELS[kernel] if isinstance(kernel, str) else int(kernel)
    cfg.integrator = INTEGRATORS[integrator] if isinstance(integrator, str) else int(integrator)
    cfg.method_xsph = int(bool(consts['useXSPH']))
    cfg.integrator_xsph = int(bool(consts['useXSPH'] if integrator_xsph is None else integrator_xsph))
    cfg.strict = int(bool(strict))
    cfg.summation_density = int(bool(consts.get('useSummationDensity', False)))
    cfg.dynamic_h = H_KEEP if keep_h else (H_DYNAMIC if fixed_h is None else H_FIXED)
    cfg.fixed_h = 0.0 if fixed_h is None else float(fixed_h)
    cfg.reorder_every = reorder_every
    return cfg


class Context:
    """One device-resident particle set + the kernels of the WCSPH step (one per GPU)."""

    def __init__(self, cfg):
        self._L = lib()
        self._h = C.c_void_p()
        self.cfg = cfg
        rc = self._L.osph_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise OsphError(rc, self._L.osph_last_error(None).decode())
        self._shape = None

    def close(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self._h = None
            self._L.osph_destroy(h)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise OsphError(rc, self._L.osph_last_error(self._h).decode())

    # ---- transfers ----
    def upload(self, pA):
        assert pA.dtype.itemsize >= 154 and pA.flags['C_CONTIGUOUS']
        self._shape = (len(pA), pA.dtype.itemsize)
        self._ck(self._L.osph_upload_aos(self._h, pA.ctypes.data, len(pA), pA.dtype.itemsize))

    def download(self, pA):
        self._ck(self._L.osph_download_aos(self._h, pA.ctypes.data, len(pA), pA.dtype.itemsize))
        return pA

    def import_device(self, ptr, n, stride=154):
        self._shape = (n, stride)
        self._ck(self._L.osph_import_device_aos(self._h, C.c_void_p(ptr), n, stride))

    def export_device(self, ptr, n, stride=154):
        self._ck(self._L.osph_export_device_aos(self._h, C.c_void_p(ptr), n, stride))

    def download_fields(self, names):
        n = self.num_active
        cols = [np.empty(n, dtype=np.float64) for _ in names]
        ids = (C.c_int32 * len(names))(*[FIELD_ID[f] for f in names])
        ptrs = (C.POINTER(C.c_double) * len(names))(*[c.ctypes.data_as(C.POINTER(C.c_double)) for c in cols])
        self._ck(self._L.osph_download_fields(self._h, len(names), ids, ptrs))
        return dict(zip(names, cols))

    def export_begin(self, names, rows=False):
        """Start an asynchronous export of the named columns; returns a ticket for export_end.
        rows=True: one value per row of the uploaded array (deleted rows keep their uploaded value)."""
        ids = (C.c_int32 * len(names))(*[FIELD_ID[f] for f in names])
        t = C.c_int64(-1)
        self._ck(self._L.osph_export_begin(self._h, len(names), ids, 1 if rows else 0, C.byref(t)))
        return (int(t.value), tuple(names), self._shape[0] if rows else self.num_active)

    def export_end(self, ticket):
        t, names, n = ticket
        cols = [np.empty(n, dtype=np.float64) for _ in names]
        ptrs = (C.POINTER(C.c_double) * len(names))(*[c.ctypes.data_as(C.POINTER(C.c_double)) for c in cols])
        self._ck(self._L.osph_export_end(self._h, t, len(names), ptrs, n))
        return dict(zip(names, cols))

    def download_rows(self, rows, pA):
        """Refresh the records pA[rows] (host row numbers of active rows) from the device; the rest of pA is untouched."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        assert pA.flags['C_CONTIGUOUS'] and (len(pA), pA.dtype.itemsize) == self._shape
        self._ck(self._L.osph_download_rows(self._h, len(rows), rows.ctypes.data_as(C.POINTER(C.c_int64)), pA.ctypes.data,
                                            pA.dtype.itemsize))
        return pA

    def upload_rows(self, rows, pA):
        """Overwrite the device state of the active rows `rows` from pA[rows]."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        assert pA.flags['C_CONTIGUOUS'] and (len(pA), pA.dtype.itemsize) == self._shape
        self._ck(self._L.osph_upload_rows(self._h, len(rows), rows.ctypes.data_as(C.POINTER(C.c_int64)), pA.ctypes.data,
                                          pA.dtype.itemsize))

    def set_active(self, active):
        """Keep the rows with active[r] != 0, mark the others deleted -- on the device (reference src/Solver.py:428-442);
        one byte per row of the uploaded array."""
        m = np.ascontiguousarray(np.asarray(active) != 0, dtype=np.uint8)
        self._ck(self._L.osph_set_active(self._h, m.ctypes.data, len(m)))

    def upload_fields(self, cols):
        names = list(cols)
        arrs = [np.ascontiguousarray(cols[f], dtype=np.float64) for f in names]
        assert all(len(a) == self.num_active for a in arrs)
        ids = (C.c_int32 * len(names))(*[FIELD_ID[f] for f in names])
        ptrs = (C.POINTER(C.c_double) * len(names))(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs])
        self._ck(self._L.osph_upload_fields(self._h, len(names), ids, ptrs))

    @property
    def num_active(self):
        return int(self._L.osph_num_active(self._h))

    @property
    def num_fluid(self):
        return int(self._L.osph_num_fluid(self._h))

    def initialize(self):
        self._ck(self._L.osph_initialize(self._h))

    # ---- the per-step calls ----
    def timestep(self):
        out = (C.c_double * 3)()
        self._ck(self._L.osph_timestep(self._h, out))
        return out[0], out[1], out[2]

    def predict(self, dt, damping):
        self._ck(self._L.osph_predict(self._h, dt, damping))

    def build_neighbours(self):
        self._ck(self._L.osph_build_neighbours(self._h))

    def compute(self):
        self._ck(self._L.osph_compute(self._h))

    def correct(self, dt, damping):
        self._ck(self._L.osph_correct(self._h, dt, damping))

    def step(self, nsteps=1, fixed_dt=None, damping=0.0):
        self._ck(self._L.osph_step(self._h, nsteps, -1.0 if fixed_dt is None else fixed_dt, damping))

    def dt_log(self, cap=65536):
        out = np.zeros((cap, 3))
        cnt = C.c_int64(0)
        self._ck(self._L.osph_get_dt_log(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(cnt)))
        return out[:min(cnt.value, cap)]

    def kinetic_energy(self):
        ke = C.c_double(0)
        self._ck(self._L.osph_kinetic_energy(self._h, C.byref(ke)))
        return ke.value

    def sort_stats(self):
        """(builds, sorts): neighbour-structure builds so far and how many of them sorted (the rest reused the last binning)."""
        out = (C.c_int64 * 2)()
        self._ck(self._L.osph_sort_stats(self._h, out))
        return int(out[0]), int(out[1])

    def sync(self):
        st = C.c_uint32(0)
        self._ck(self._L.osph_sync(self._h, C.byref(st)))
        return st.value

    # ---- validation / queries ----
    def cells(self):
        grid = (C.c_double * 7)()
        ids = np.zeros(self.num_active, dtype=np.int64)
        self._ck(self._L.osph_get_cells(self._h, grid, ids.ctypes.data_as(C.POINTER(C.c_int64))))
        g = dict(xmin=grid[0], xmax=grid[1], ymin=grid[2], ymax=grid[3], cell_size=grid[4],
                 ncx=int(grid[5]), ncy=int(grid[6]))
        return g, ids

    def neighbours_csr(self):
        n = self.num_active
        off = np.zeros(n + 1, dtype=np.int64)
        total = C.c_int64(0)
        ip = C.POINTER(C.c_int64)
        self._ck(self._L.osph_get_neighbours_csr(self._h, off.ctypes.data_as(ip), None, 0, C.byref(total)))
        idx = np.zeros(max(total.value, 1), dtype=np.int64)
        self._ck(self._L.osph_get_neighbours_csr(self._h, off.ctypes.data_as(ip), idx.ctypes.data_as(ip),
                                                 len(idx), C.byref(total)))
        return off, idx[:total.value]

    def near_pos(self, x, y, h, cap=256):
        ip, dp = C.POINTER(C.c_int64), C.POINTER(C.c_double)
        while True:
            idx = np.zeros(cap, dtype=np.int64); r = np.zeros(cap); q = np.zeros(cap); hh = np.zeros(cap)
            cnt = C.c_int64(0)
            self._ck(self._L.osph_near_pos(self._h, x, y, h, cap, idx.ctypes.data_as(ip), r.ctypes.data_as(dp),
                                           q.ctypes.data_as(dp), hh.ctypes.data_as(dp), C.byref(cnt)))
            if cnt.value <= cap:
                k = cnt.value
                return hh[:k], q[:k], r[:k], idx[:k]
            cap = int(cnt.value)

    # ---- slab decomposition (device pointers of caller-owned buffers) ----
    def reserve(self, capacity):
        self._ck(self._L.osph_reserve(self._h, int(capacity)))

    def set_row_ids(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._ck(self._L.osph_set_row_ids(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), len(ids)))

    def slab_configure(self, x_lo, x_hi, ghost_ptr, ghost_capacity):
        self._ck(self._L.osph_slab_configure(self._h, x_lo, x_hi, C.c_void_p(ghost_ptr), ghost_capacity))

    def slab_step_plan(self, step, nsteps):
        self._ck(self._L.osph_slab_step_plan(self._h, step, nsteps))

    def slab_dt_local(self, out_ptr):
        self._ck(self._L.osph_slab_dt_local(self._h, C.c_void_p(out_ptr)))

    def slab_step_begin(self, dt_ptr, fixed_dt, damping):
        self._ck(self._L.osph_slab_step_begin(self._h, C.c_void_p(dt_ptr) if dt_ptr else None,
                                              -1.0 if fixed_dt is None else fixed_dt, damping))

    def slab_pack(self, width, mig_l, mig_r, mig_cap, halo_l, halo_r, halo_cap, meta_ptr):
        self._ck(self._L.osph_slab_pack(self._h, width, C.c_void_p(mig_l), C.c_void_p(mig_r), mig_cap,
                                        C.c_void_p(halo_l), C.c_void_p(halo_r), halo_cap, C.c_void_p(meta_ptr)))

    def slab_commit(self, n_mig_out, mig_in_ptr, n_mig_in, n_ghost, bounds):
        b = (C.c_double * 6)(*bounds)
        self._ck(self._L.osph_slab_commit(self._h, n_mig_out, C.c_void_p(mig_in_ptr) if mig_in_ptr else None,
                                          n_mig_in, n_ghost, b))

    def slab_step_end(self, damping):
        self._ck(self._L.osph_slab_step_end(self._h, damping))

    def download_owned(self, pA, ids):
        n = C.c_int64(0)
        self._ck(self._L.osph_download_owned(self._h, pA.ctypes.data, len(pA), pA.dtype.itemsize,
                                             ids.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
        return n.value

    def slab_export(self, ids_ptr, label_ptr, fields, col_ptrs):
        ids = (C.c_int32 * len(fields))(*[FIELD_ID[f] for f in fields])
        ptrs = (C.c_void_p * len(fields))(*col_ptrs)
        self._ck(self._L.osph_slab_export(self._h, C.c_void_p(ids_ptr), C.c_void_p(label_ptr) if label_ptr else None,
                                          len(fields), ids, ptrs))

    # ---- slab decomposition sequenced in the library (direct NCCL) ----
    def slab_comm_create(self, unique_id, rank, world, x_lo, x_hi, r0, hmax, mig_cap, halo_cap):
        comm = C.c_void_p()
        self._ck(self._L.osph_slab_comm_create(self._h, nccl_library_path().encode(), bytes(unique_id), rank, world,
                                               x_lo, x_hi, r0, hmax, mig_cap, halo_cap, C.byref(comm)))
        return comm

    def slab_comm_destroy(self, comm):
        self._L.osph_slab_comm_destroy(self._h, comm)

    def slab_comm_attach(self, comm):
        self._ck(self._L.osph_slab_comm_attach(self._h, comm))

    def slab_comm_set_bounds(self, comm, x_lo, x_hi):
        self._ck(self._L.osph_slab_comm_set_bounds(self._h, comm, x_lo, x_hi))

    def slab_run(self, comm, nsteps, fixed_dt=None, damping=0.0):
        self._ck(self._L.osph_slab_run(self._h, comm, nsteps, -1.0 if fixed_dt is None else fixed_dt, damping))

    def slab_last_counts(self, comm):
        out = (C.c_int64 * 8)()
        self._L.osph_slab_last_counts(comm, out)
        return list(out)

    def probe_pressure(self, x, y, h):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        rho = np.empty_like(x); p = np.empty_like(x)
        dp = C.POINTER(C.c_double)
        self._ck(self._L.osph_probe_pressure(self._h, x.size, x.ctypes.data_as(dp), y.ctypes.data_as(dp), float(h),
                                             rho.ctypes.data_as(dp), p.ctypes.data_as(dp)))
        return rho, p

    # ---- slab decomposition over NVLink peer memory (CUDA IPC windows, mailbox kernels) ----
    def slab_p2p_create(self, rank, world, x_lo, x_hi, r0, hmax, mig_cap, halo_cap):
        h = C.c_void_p(); buf = C.create_string_buffer(64)
        self._ck(self._L.osph_slab_p2p_create(self._h, rank, world, x_lo, x_hi, r0, hmax, mig_cap, halo_cap, C.byref(h), buf))
        return h, buf.raw

    def slab_p2p_connect(self, p2p, all_handles):
        self._ck(self._L.osph_slab_p2p_connect(self._h, p2p, bytes(all_handles)))

    def slab_p2p_destroy(self, p2p):
        self._L.osph_slab_p2p_destroy(self._h, p2p)

    def slab_p2p_attach(self, p2p):
        self._ck(self._L.osph_slab_p2p_attach(self._h, p2p))

    def slab_p2p_set_bounds(self, p2p, x_lo, x_hi):
        self._ck(self._L.osph_slab_p2p_set_bounds(self._h, p2p, x_lo, x_hi))

    def slab_p2p_run(self, p2p, nsteps, fixed_dt=None, damping=0.0):
        self._ck(self._L.osph_slab_p2p_run(self._h, p2p, nsteps, -1.0 if fixed_dt is None else fixed_dt, damping))

    def slab_p2p_last_counts(self, p2p):
        out = (C.c_int64 * 8)()
        self._L.osph_slab_p2p_last_counts(p2p, out)
        return list(out)

    def slab_p2p_stats(self, p2p):
        """(sorting steps, reusing steps) of the peer-memory sequencer's slab cadence."""
        out = (C.c_int64 * 2)()
        self._ck(self._L.osph_slab_p2p_stats(p2p, out))
        return int(out[0]), int(out[1])

    def timers(self):
        out = (C.c_double * 6)()
        self._ck(self._L.osph_get_timers(self._h, out))
        return dict(zip(('time_step', 'integrate_prediction', 'neighbour_hood', 'compute',
                         'integrate_correction', 'transfer'), [v * 1e-3 for v in out]))

    @property
    def launch_count(self):
        return int(self._L.osph_launch_count(self._h))

    @property
    def stream(self):
        return int(self._L.osph_stream(self._h))

    def pair_kernel_info(self):
        """(launches of the fused pair kernel so far, those that ran its uniform-smoothing-length instantiation, build
        flags: 1 that instantiation exists, 2 sign-bit clamps, 4 software-pipelined flush loop)."""
        out = (C.c_int64 * 3)()
        self._ck(self._L.osph_pair_kernel_info(self._h, out))
        return int(out[0]), int(out[1]), int(out[2])

    def pair_kernel_time(self):
        us = C.c_double(0); n = C.c_int64(0)
        self._ck(self._L.osph_pair_kernel_time(self._h, C.byref(us), C.byref(n)))
        return us.value, n.value


# ---- stand-alone leaf functions (host arrays in, host arrays out; computed on the device) ----------

def _leaf_ck(rc):
    if rc != 0:
        raise OsphError(rc, lib().osph_leaf_last_error().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def nccl_library_path():
    """Fallback path of libnccl.so.2 (the copy bundled with torch); the library prefers one already loaded."""
    try:
        import nvidia.nccl
        p = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
        return p if os.path.exists(p) else ""
    except Exception:
        return ""


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().osph_nccl_unique_id(nccl_library_path().encode(), buf)
    if rc != 0:
        raise OsphError(rc, lib().osph_nccl_last_error().decode())
    return buf.raw


def default_device():
    return int(os.environ.get("OSPH_DEVICE", "0"))


def precision_from_env():
    """OSPH_PRECISION=fp64 (default, validation mode) | fp32 (performance mode): selects the pair-kernel arithmetic
    without editing example scripts."""
    v = os.environ.get("OSPH_PRECISION", "fp64").lower()
    if v not in ("fp64", "fp32"):
        raise ValueError("OSPH_PRECISION must be fp64 or fp32")
    return FP64 if v == "fp64" else FP32


def leaf_kernel(kernel, what, x, r, h):
    r = _f64(r); h = _f64(np.broadcast_to(h, r.shape)); out = np.empty_like(r)
    x = _f64(np.broadcast_to(x, r.shape)) if x is not None else None
    _leaf_ck(lib().osph_leaf_kernel(default_device(), KERNELS[kernel], what, r.size,
                                    _ptr(x) if x is not None else None, _ptr(r), _ptr(h), _ptr(out)))
    return out


def leaf_tait_pressure(rho, label, gamma, B, rho0, Pb=0.0):
    rho = _f64(rho); lab = np.ascontiguousarray(np.broadcast_to(label, rho.shape), dtype=np.int8)
    out = np.empty_like(rho)
    _leaf_ck(lib().osph_leaf_tait_pressure(default_device(), rho.size, _ptr(rho),
                                           lab.ctypes.data_as(C.POINTER(C.c_int8)), gamma, B, rho0, Pb, _ptr(out)))
    return out


def leaf_tait_height(y, rho0, H, B, gamma):
    y = _f64(y); out = np.empty_like(y)
    _leaf_ck(lib().osph_leaf_tait_height(default_device(), y.size, _ptr(y), rho0, H, B, gamma, _ptr(out)))
    return out


def leaf_compute_h(sigma, m, rho):
    m = _f64(m); rho = _f64(rho); out = np.empty_like(m)
    _leaf_ck(lib().osph_leaf_compute_h(default_device(), m.size, sigma, _ptr(m), _ptr(rho), _ptr(out)))
    return out


#: column order of a computed-neighbour table crossing the C ABI (include/osph.h: OSPH_COMP_*)
COMP_COLUMNS = ['m', 'p', 'rho', 'h', 'c', 'r', 'w', 'dw_x', 'dw_y', 'x', 'y', 'vx', 'vy']


def _self_field(p, name):
    """Field of the self particle; the reference passes a particle_dtype record, some of its tests an empty array."""
    try:
        return float(p[name])
    except (KeyError, IndexError, ValueError, TypeError):
        return 0.0


def leaf_equations(p, comp, alpha=0.0, beta=0.0, epsilon=0.0, r0=0.0, D=0.0, p1=4.0, p2=2.0):
    """Continuity, Momentum, XSPH and BoundaryForce of particle record `p` over the computed-neighbour table `comp`
    (computed_dtype) in one device launch.  Returns dict(drho, a=[ax, ay], xsph=[x, y], f=[fx, fy])."""
    comp = np.atleast_1d(comp)
    J = len(comp)
    cols = np.empty((len(COMP_COLUMNS), J), dtype=np.float64)
    for k, f in enumerate(COMP_COLUMNS):
        cols[k] = comp[f]
    lab = np.ascontiguousarray(comp['label'], dtype=np.int8)
    out = np.zeros(7)
    rho_self = _self_field(p, 'rho')
    _leaf_ck(lib().osph_leaf_equations(default_device(), J, lab.ctypes.data_as(C.POINTER(C.c_int8)), _ptr(cols),
                                       _self_field(p, 'p'), rho_self, _self_field(p, 'h'), _self_field(p, 'c'),
                                       float(alpha), float(beta), float(epsilon), float(r0), float(D), float(p1), float(p2),
                                       _ptr(out)))
    return dict(drho=float(out[0]), a=[float(out[1]), float(out[2])], xsph=[float(out[3]), float(out[4])],
                f=[float(out[5]), float(out[6])])


def leaf_courant(alpha, h, c):
    h = _f64(np.atleast_1d(h)); c = _f64(np.atleast_1d(c))
    if h.size != c.size:
        raise ValueError("h and c must have the same length")
    out = np.zeros(1)
    _leaf_ck(lib().osph_leaf_courant(default_device(), float(alpha), h.size, _ptr(h), _ptr(c), _ptr(out)))
    return float(out[0])


def leaf_differences(self4, nbr):
    """(x, y, vx, vy) of the self particle minus the same four columns of its neighbours; nbr: array (4, J)."""
    nbr = _f64(nbr)
    out = np.empty_like(nbr)
    s4 = _f64(self4)
    _leaf_ck(lib().osph_leaf_differences(default_device(), nbr.shape[1], _ptr(s4), _ptr(nbr), _ptr(out)))
    return out
