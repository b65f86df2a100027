"""Particle-lattice generator with the call signature of reference src/Helpers.py:7-87."""
from math import ceil

import numpy as np

from src.Common import particle_dtype, ParticleType


class Helpers:
    @staticmethod
    def rect(xmin: float, xmax: float, ymin: float, ymax: float, r0: float, mass: float = 1.0,
             rho0: float = 1000.0, pack: bool = False, label: ParticleType = ParticleType.Fluid,
             strict: bool = False, packDirection: str = 'y') -> np.array:
        """Rectangular block of particles, optionally hex-packed.

        Nx = max(1, ceil((xmax-xmin)/r0)) points are spread with np.linspace over
        [xmin, xmax] (same for y), x-major ordering.  With ``pack`` every other
        column (packDirection 'y') or the columns picked by the reference's
        ``2*i::Nx`` stride (packDirection 'x', reference Helpers.py:71-74) is
        shifted by r0/2; ``strict`` then drops particles pushed outside the box.
        """
        xspan = xmax - xmin
        yspan = ymax - ymin
        Nx = max(1, ceil(xspan / r0))
        Ny = max(1, ceil(yspan / r0))

        gx, gy = np.meshgrid(np.linspace(xmin, xmax, Nx), np.linspace(ymin, ymax, Ny), indexing='ij')
        pA = np.zeros(Nx * Ny, dtype=particle_dtype)
        pA['label'] = label
        pA['x'] = gx.ravel()
        pA['y'] = gy.ravel()
        pA['m'] = mass
        pA['rho'] = rho0

        if pack and xspan > 0:
            if packDirection == 'y':
                shift = np.zeros((Nx, Ny))
                shift[0:2 * (Nx // 2):2, :] = r0 / 2
                pA['y'] += shift.ravel()
            elif packDirection == 'x':
                for i in range(Ny // 2):
                    pA['x'][2 * i::Nx] += r0 / 2
            else:
                raise Exception('Invalid packing direction.')

            if strict:
                keep = (pA['x'] <= xmax) & (pA['x'] >= xmin) & (pA['y'] <= ymax) & (pA['y'] >= ymin)
                pA = pA[keep]
        return pA
