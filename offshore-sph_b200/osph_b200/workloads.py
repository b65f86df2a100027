"""Synthetic particle sets for tests and benchmarks (host side, setup time only).

The geometries restate what the reference's example scripts build
(examples/DamBreak.py:16-55, examples/Containment.py:15-44,
examples/IceBreak.py:23-78) using the drop-in `src.Helpers.rect`, plus the
`Solver.setup()` initialisation (src/Solver.py:173-208) so that the arrays are
ready for the first time step.  SURVEY.md section 8(d) fixes the perturbation
used before timing: a perfect lattice at rest has vanishing pair forces.
"""
import math

import numpy as np

from src.Common import ParticleType, particle_dtype
from src.Helpers import Helpers


def wcsph_constants(height, r0, rho0=1000.0, useXSPH=True, Pb=0.0, useSummationDensity=False):
    """The constants WCSPH.__init__ derives (reference src/Methods/WCSPH.py:56-79)."""
    co = 10.0 * math.sqrt(2 * 9.81 * height)
    gamma = 7.0
    return dict(height=float(height), r0=float(r0), rho0=float(rho0), Pb=float(Pb), gamma=gamma, co=co,
                B=co * co * rho0 / gamma, alpha=0.01, beta=0.0, epsilon=0.5,
                D=5 * 9.81 * height, p1=4.0, p2=2.0, useXSPH=bool(useXSPH),
                useSummationDensity=bool(useSummationDensity))


def hydrostatic_density(y, height, rho0, B, gamma):
    """rho0 * (1 + rho0 g (H - y) / B)^(1/gamma), reference src/Equations/TaitEOS.py:46-65."""
    return rho0 * (1 + rho0 * 9.81 * (height - y) / B) ** (1 / gamma)


def tait_pressure(rho, label, rho0, B, gamma, Pb=0.0):
    """Tait EOS, fluid rows only (reference src/Equations/TaitEOS.py:6-31)."""
    p = ((rho / rho0) ** gamma - 1.0) * B
    return np.where(label == ParticleType.Fluid, p, 0.0) + Pb


def setup_state(pA, consts, h=None):
    """What Solver.setup() does to the fluid rows (reference src/Solver.py:184-196)."""
    f = (pA['label'] == ParticleType.Fluid) & (~pA['deleted'])
    if h is None:
        hh = np.zeros(f.sum())
        ok = pA['rho'][f] > 1e-12
        hh[ok] = 1.3 * (pA['m'][f][ok] / pA['rho'][f][ok]) ** 0.5
        pA['h'][f] = hh
    else:
        pA['h'][f] = h
    pA['rho'][f] = hydrostatic_density(pA['y'][f], consts['height'], consts['rho0'], consts['B'], consts['gamma'])
    pA['p'][f] = tait_pressure(pA['rho'][f], pA['label'][f], consts['rho0'], consts['B'], consts['gamma'], consts['Pb'])
    pA['c'][f] = consts['co']
    return pA


def perturb(pA, r0, height, seed=0, pos=0.05, vel=0.1):
    """Deterministic jitter of the fluid rows."""
    rng = np.random.default_rng(seed)
    f = np.flatnonzero((pA['label'] == ParticleType.Fluid) & (~pA['deleted']))
    u = rng.uniform(-1.0, 1.0, size=(4, len(f)))
    pA['x'][f] += pos * r0 * u[0]
    pA['y'][f] += pos * r0 * u[1]
    v = vel * math.sqrt(9.81 * height)
    pA['vx'][f] = v * u[2]
    pA['vy'][f] = v * u[3]
    return pA


def dam_break(N, rho0=1000.0, height=25.0, length=180.0, wall=30.0, temp_wall=True):
    """Dam-break tank: hex-packed fluid block [0,25]^2, floor, left wall, removable gate.

    N fluid particles per side; r0 = 25/N; mass = (25/N)^2 rho0.
    Returns (r0, pA) like the example's create_particles.
    """
    r0 = 25.0 / N
    mass = 25.0 * 25.0 / N ** 2 * rho0
    lo = -r0
    parts = [
        Helpers.rect(xmin=0, xmax=25, ymin=0, ymax=25, r0=r0, mass=mass, rho0=rho0, pack=True),
        Helpers.rect(xmin=lo, xmax=length, ymin=lo, ymax=lo, r0=r0, mass=0., rho0=0., label=ParticleType.Boundary),
        Helpers.rect(xmin=lo, xmax=lo, ymin=lo, ymax=wall, r0=r0, mass=0., rho0=0., label=ParticleType.Boundary),
    ]
    if temp_wall:
        parts.append(Helpers.rect(xmin=25 - lo, xmax=25 - lo, ymin=lo, ymax=wall, r0=r0, mass=0., rho0=0.,
                                  label=ParticleType.TempBoundary))
    return r0, np.concatenate(parts)


def dam_break_case(N, hfac=1.6, seed=0, perturbed=True, **kw):
    """Dam break of `dam_break(N)` after setup (+ jitter): returns dict(pA, consts, r0, h)."""
    r0, pA = dam_break(N, **kw)
    consts = wcsph_constants(25.0, r0, 1000.0, True, 0.0, False)
    h = hfac * r0
    setup_state(pA, consts, h)
    if perturbed:
        perturb(pA, r0, 25.0, seed)
    return dict(pA=pA, consts=consts, r0=r0, h=h)


def tank(Nx, width=1.0, height=1.0, rho0=1000.0, pack=True, lid=1.3):
    """Closed tank (floor + two walls) filled to `height`: the Containment geometry."""
    r0 = width / Nx
    fluid = Helpers.rect(xmin=0, xmax=width, ymin=0, ymax=height, r0=r0, mass=1.0, rho0=rho0, pack=pack)
    fluid['m'] = width * height / len(fluid) * rho0
    lo, hi_x, hi_y = -r0, width + r0, lid * height + r0
    bnd = dict(r0=r0, mass=0., rho0=0., label=ParticleType.Boundary)
    return r0, np.concatenate((
        fluid,
        Helpers.rect(xmin=lo, xmax=hi_x, ymin=lo, ymax=lo, **bnd),
        Helpers.rect(xmin=lo, xmax=lo, ymin=lo, ymax=hi_y, **bnd),
        Helpers.rect(xmin=hi_x, xmax=hi_x, ymin=lo, ymax=hi_y, **bnd)))


def tank_case(Nx, width=1.0, height=1.0, h=None, useXSPH=False, seed=0, perturbed=True, coupled_row=False):
    """Tank after setup; `coupled_row` adds a floating row of Coupled particles (IceBreak-like)."""
    r0, pA = tank(Nx, width, height)
    if coupled_row:
        f = pA[pA['label'] == ParticleType.Fluid]
        ice = Helpers.rect(xmin=-r0, xmax=0.6 * width, ymin=f['y'].max() + r0, ymax=f['y'].max() + r0, r0=r0,
                           mass=f['m'][0], rho0=1000.0, label=ParticleType.Coupled)
        pA = np.concatenate((pA, ice))
    consts = wcsph_constants(height, r0, 1000.0, useXSPH, 0.0, False)
    setup_state(pA, consts, h)
    if perturbed:
        perturb(pA, r0, height, seed)
    return dict(pA=pA, consts=consts, r0=r0, h=h)
