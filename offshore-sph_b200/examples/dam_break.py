"""2-D dam break on the B200 path (same set-up as the reference's examples/DamBreak.py: Wendland kernel, PEC,
XSPH, fixed h = 1.6 r0).  The reference's own example scripts also run unedited against this package when
placed next to `src/`; this script is the repository's own, parameterised from the command line.

    python examples/dam_break.py --n 200 --duration 0.5 [--precision fp32]
"""
import argparse
import os
import sys

sys.path.insert(1, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from src.Solver import Solver                      # noqa: E402
from src.Methods.WCSPH import WCSPH                # noqa: E402
from src.Kernels.Wendland import Wendland          # noqa: E402
from src.Kernels.CubicSpline import CubicSpline    # noqa: E402
from src.Integrators.PEC import PEC                # noqa: E402
from osph_b200 import workloads                    # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=50, help='fluid particles per side')
    ap.add_argument('--duration', type=float, default=1.0)
    ap.add_argument('--max-settle', type=int, default=500)
    ap.add_argument('--kernel', default='wendland', choices=['wendland', 'cubic'])
    ap.add_argument('--precision', default=None, choices=['fp64', 'fp32'])
    ap.add_argument('--out', default=None)
    a = ap.parse_args(argv)
    if a.precision:
        os.environ['OSPH_PRECISION'] = a.precision

    r0, pA = workloads.dam_break(a.n)
    kernel = Wendland() if a.kernel == 'wendland' else CubicSpline()
    method = WCSPH(height=25.0, r0=r0, rho0=1000.0, useXSPH=True, Pb=0, useSummationDensity=False)
    solver = Solver(method, PEC(useXSPH=True, strict=False), kernel, a.duration, incrementalWriteout=False,
                    h=1.6 * r0, maxSettle=a.max_settle)
    solver.addParticles(pA)
    solver.setup()
    solver.run()
    solver.timing()
    if a.out:
        solver.save(a.out)
    return solver


if __name__ == '__main__':
    main()
