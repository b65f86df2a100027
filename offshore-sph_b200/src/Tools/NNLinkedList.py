"""Neighbour search with the interface of reference src/Tools/NNLinkedList.py:21-84.

update(pA) uploads the (active) particle array and builds the device structure: strict-IEEE cell keys on the
reference grid, radix sort by key, cell table (csrc/step.cu, csrc/sort.cu).  near / nearPos run the reference
predicate (3x3 reference cells AND r/h_ij <= 3.0) on the device.  Neighbours come back sorted by index; the
reference returns the same SET in cell-walk order.
"""
import numpy as np


class NNLinkedList:
    def __init__(self, scale: float = 2.0):
        self.scale = scale
        self.shifts = np.array([-1, 0, 1], dtype=np.int64)
        self._ctx = None
        self._own = False
        self.xmin = self.xmax = self.ymin = self.ymax = 0.0
        self.cell_size = 0.0
        self.n_cells = 0
        self.ncells_per_dim = np.array([0, 0], dtype=np.int64)

    # -- binding to a Solver's context (Solver.run keeps one resident particle set) --
    def _bind(self, ctx):
        self._release()
        self._ctx, self._own = ctx, False

    def _release(self):
        if self._own and self._ctx is not None:
            self._ctx.close()
        self._ctx, self._own = None, False

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _refresh_params(self):
        g, _ = self._ctx.cells()
        self.xmin, self.xmax, self.ymin, self.ymax = g['xmin'], g['xmax'], g['ymin'], g['ymax']
        self.cell_size = g['cell_size']
        self.ncells_per_dim = np.array([g['ncx'], g['ncy']], dtype=np.int64)
        self.n_cells = int(g['ncx'] * g['ncy'])

    def update(self, pA: np.array):
        """Stand-alone use: (re)build from a host array."""
        from osph_b200 import capi
        from src.Common import particle_dtype
        if self._ctx is None or not self._own:
            consts = dict(height=1.0, r0=1.0, rho0=1000.0, Pb=0.0, gamma=7.0, co=1.0, B=1.0, alpha=0.0, beta=0.0,
                          epsilon=0.5, D=0.0, p1=4.0, p2=2.0, useXSPH=False)
            cfg = capi.make_config(consts, 'cubic', 'pec', capi.FP64, None, keep_h=True, device=capi.default_device())
            cfg.nn_scale = self.scale
            self._ctx, self._own = capi.Context(cfg), True
        arr = np.ascontiguousarray(pA).astype(particle_dtype, copy=True)
        arr['deleted'] = False
        self._ctx.upload(arr)
        self._ctx.build_neighbours()
        self._refresh_params()

    def nearPos(self, x: float, y: float, h: float, pA: np.array = None):
        hh, q, r, idx = self._ctx.near_pos(float(x), float(y), float(h))
        return hh, q, r, idx.astype(np.uint64)

    def near(self, i: int, pA: np.array):
        return self.nearPos(pA[i]['x'], pA[i]['y'], pA[i]['h'], pA)
