"""Closed tank at rest (the reference's examples/Containment.py set-up: cubic spline, PEC without XSPH,
dynamic smoothing length h = 1.3 sqrt(m/rho)).

    python examples/containment.py --nx 60 --duration 0.2
"""
import argparse
import os
import sys

sys.path.insert(1, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from src.Solver import Solver                      # noqa: E402
from src.Methods.WCSPH import WCSPH                # noqa: E402
from src.Kernels.CubicSpline import CubicSpline    # noqa: E402
from src.Integrators.PEC import PEC                # noqa: E402
from osph_b200 import workloads                    # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=30)
    ap.add_argument('--duration', type=float, default=0.5)
    ap.add_argument('--max-settle', type=int, default=500)
    ap.add_argument('--out', default=None)
    a = ap.parse_args(argv)

    r0, pA = workloads.tank(a.nx, 1.0, 1.0)
    method = WCSPH(height=1.0, rho0=1000.0, r0=r0, useXSPH=False, Pb=0)      # no useSummationDensity: as the reference example
    solver = Solver(method, PEC(useXSPH=False, strict=False), CubicSpline(), a.duration, quick=False,
                    incrementalWriteout=False, exportProperties=['x', 'y', 'p', 'vx', 'vy'], maxSettle=a.max_settle)
    solver.addParticles(pA)
    solver.setup()
    solver.run()
    solver.timing()
    if a.out:
        solver.save(a.out)
    return solver


if __name__ == '__main__':
    main()
