"""Import the UNMODIFIED reference (read-only at /root/reference) under numba >= 0.59.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the reference
tree does not travel to the GPU box, so nothing that runs there may import this.
It is used by tests/golden/make_golden.py to freeze golden vectors and by the
container-only cross-check tests (skipped when /root/reference is absent).

What the shim does: restores `numba.jitclass`, stubs the
three absent display/IO packages that src/Solver.py imports, and swaps the
object-mode `findActive` for its pure-Python body.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OSPH_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _isolate_path():
    """Put the reference first and hide every other directory that provides a `src` package."""
    saved = list(sys.path)
    keep = [p for p in saved if not os.path.isdir(os.path.join(p or ".", "src"))]
    sys.path[:] = [REFERENCE_ROOT] + keep
    return saved


def load():
    """Return a namespace with the reference's hot-path classes and functions."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import numba
    from numba.experimental import jitclass
    numba.jitclass = jitclass
    for name in ("h5py", "colorama", "prettytable"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["colorama"], "Fore"):
        sys.modules["colorama"].Fore = types.SimpleNamespace(YELLOW="", GREEN="", RED="")
        sys.modules["colorama"].Style = types.SimpleNamespace(RESET_ALL="")
    if not hasattr(sys.modules["prettytable"], "PrettyTable"):
        sys.modules["prettytable"].PrettyTable = object

    # Make `src` resolve to the reference, whatever was imported before.
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    saved_path = _isolate_path()
    try:
        import src.Tools.SolverTools as ST
        ST.findActive = ST.findActive.py_func
        import src.Solver as S
        S.findActive = ST.findActive
        from src.Common import particle_dtype, computed_dtype, ParticleType
        from src.Tools.NNLinkedList import NNLinkedList
        from src.Methods.WCSPH import WCSPH
        from src.Integrators.PEC import PEC
        from src.Integrators.Verlet import Verlet
        from src.Equations.TimeStep import TimeStep
        from src.Equations.KineticEnergy import KineticEnergy
        from src.Equations.Continuity import Continuity
        from src.Equations.Momentum import Momentum
        from src.Equations.XSPH import XSPH
        from src.Equations.BoundaryForce import BoundaryForce
        from src.Equations.TaitEOS import TaitEOS, TaitEOS_B, TaitEOS_co, TaitEOS_height
        from src.Kernels.CubicSpline import CubicSpline
        from src.Kernels.Wendland import Wendland
        from src.Kernels.Gaussian import Gaussian
        from src.Helpers import Helpers
        ns = types.SimpleNamespace(
            ST=ST, Solver=S.Solver, particle_dtype=particle_dtype, computed_dtype=computed_dtype,
            ParticleType=ParticleType, NNLinkedList=NNLinkedList, WCSPH=WCSPH, PEC=PEC, Verlet=Verlet,
            TimeStep=TimeStep, KineticEnergy=KineticEnergy, Continuity=Continuity, Momentum=Momentum,
            XSPH=XSPH, BoundaryForce=BoundaryForce, TaitEOS=TaitEOS, TaitEOS_B=TaitEOS_B,
            TaitEOS_co=TaitEOS_co, TaitEOS_height=TaitEOS_height, CubicSpline=CubicSpline,
            Wendland=Wendland, Gaussian=Gaussian, Helpers=Helpers, _loop=ST._loop,
            computeH=ST.computeH, _assignProps=ST._assignProps)
        ns.example = _example_loader
    finally:
        sys.path[:] = saved_path
        # Leave the reference's modules registered under a private prefix only.
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            sys.modules["_osph_reference." + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return ns


def _example_loader(name):
    """Load examples/<name>.py of the reference as a module without running main()."""
    path = os.path.join(REFERENCE_ROOT, "examples", name + ".py")
    spec = importlib.util.spec_from_file_location("_osph_reference_example_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    for k in [k for k in sys.modules if k.startswith("_osph_reference.src")]:
        sys.modules[k[len("_osph_reference."):]] = sys.modules[k]
    if "src.Post.Plot" not in sys.modules:      # pyqtgraph/Qt are absent; plotting is never exercised
        stub = types.ModuleType("src.Post.Plot")
        stub.Plot = object
        sys.modules["src.Post.Plot"] = stub
    saved_path = _isolate_path()
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mod
