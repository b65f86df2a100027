"""Host interchange data model of the drop-in package.

Mirrors the public names of reference src/Common.py:8-98 (ParticleType,
get_label_code, particle_dtype, computed_dtype).  The packed 154-byte
``particle_dtype`` record is the format crossing the C ABI
(include/osph.h: osph_upload_aos / osph_download_aos); on the device the
state lives as SoA columns ("Data layout in HBM").
"""
import enum

import numpy as np


class ParticleType(enum.IntEnum):
    Fluid = 0
    Boundary = 1
    TempBoundary = 2
    Coupled = 3


_LABELS = {'fluid': ParticleType.Fluid, 'boundary': ParticleType.Boundary,
           'temp-boundary': ParticleType.TempBoundary, 'coupled': ParticleType.Coupled}


def get_label_code(label: str):
    try:
        return _LABELS[label]
    except KeyError:
        raise Exception('Argument out of range')


#: real-valued columns in record order; index in this list == field id of the C ABI (OSPH_F_*)
REAL_FIELDS = ['m', 'rho', 'p', 'c', 'drho', 'h', 'x', 'y', 'vx', 'vy', 'ax', 'ay',
               'xsphx', 'xsphy', 'x0', 'y0', 'vx0', 'vy0', 'rho0']

particle_dtype = np.dtype({
    'names': ['deleted', 'label'] + REAL_FIELDS,
    'formats': [np.bool_, np.int8] + [np.double] * len(REAL_FIELDS),
})
assert particle_dtype.itemsize == 154

# Per-neighbour scratch record of the reference's CPU loop.  The device path never
# materialises it (the pair kernel is fused); kept because leaf-equation callers and
# tests build arrays of it.
computed_dtype = np.dtype({
    'names': ['label', 'm', 'p', 'rho', 'h', 'q', 'c', 'r', 'w', 'dw_x', 'dw_y', 'x', 'y', 'vx', 'vy'],
    'formats': [np.int8] + [np.double] * 14,
})
assert computed_dtype.itemsize == 113


def _stack(m1, m2):
    """Column-stack two 1-D arrays into shape (n, 2)."""
    return np.stack((np.asarray(m1), np.asarray(m2)), axis=1)
