"""CPU: the kernel SOURCES of offshore-sph_b200/csrc executed under the SIMT emulator of tests/emu (g++ build of the
same .cu files: fibers for CUDA threads, __syncthreads / warp collectives as switch points) and held to the same
parity tests as the GPU build -- cell ids and neighbour sets bit-exact, FP64 fields within 1e-10 of the golden
vectors and the oracle, fused step loop identical to the explicit calls, and so on.

This is test infrastructure: it checks the kernels' logic (indexing, staging, warp-synchronous control flow, edge
paths) in a container without a GPU.  It is NOT a product path -- nothing under offshore-sph_b200/ knows about the
emulated library, and the GPU tests (-m gpu) remain the parity gate on the real hardware.
"""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build as emu_build  # noqa: E402

if not emu_build.available():
    pytest.skip("g++ or the CUDA headers are missing: cannot build the emulated library", allow_module_level=True)


@pytest.fixture(autouse=True, scope="module")
def _emulated_library():
    from osph_b200 import capi
    path = emu_build.build()
    saved = (capi.LIB_PATH, capi._lib)
    capi.LIB_PATH, capi._lib = path, None
    yield
    capi.LIB_PATH, capi._lib = saved


# the GPU parity tests, re-collected here without their `gpu` mark (it is attached to the modules, not the functions)
from test_gpu_parity import *  # noqa: F401,F403,E402
from test_gpu_solver import *  # noqa: F401,F403,E402
from test_gpu_edges import *  # noqa: F401,F403,E402
import test_gpu_edges as _edges  # noqa: E402
import test_gpu_parity as _parity  # noqa: E402
import test_gpu_solver as _solver  # noqa: E402


def test_pair_kernel_batched_staging_path():
    """A build with PAIR_CAP = 96 staged records: the candidate runs of a CTA never fit in shared memory together, so
    every CTA takes the batched path that production sizes only reach with very dense cells or domain-spanning rows."""
    from osph_b200 import capi
    path = emu_build.build(defines=("PAIR_CAP=96",), tag="_cap96")
    saved = (capi.LIB_PATH, capi._lib)
    capi.LIB_PATH, capi._lib = path, None
    try:
        _parity.test_dam_break_vs_oracle(60, 'wendland')
        _parity.test_dam_break_vs_oracle(150, 'gaussian')
        _parity.test_whole_steps_vs_golden('tank30_cubic_dynh')
        _edges.test_cluster_denser_than_the_candidate_list()
        _edges.test_coincident_particles_follow_the_reference_guards()
    finally:
        capi.LIB_PATH, capi._lib = saved


def test_pair_kernel_128_thread_variant():
    """-DOSPH_PAIR_THREADS=128 -DPAIR_CAP=512 (the `t128` build of tools/build_round2_variants.sh): four smaller CTAs per SM
    instead of two; list stride, staging capacity and the CTA-wide run unions all change with the CTA size."""
    from osph_b200 import capi
    path = emu_build.build(defines=("OSPH_PAIR_THREADS=128", "PAIR_CAP=512", "PAIR_MINB64=4", "PAIR_MINB32=6"), tag="_t128")
    saved = (capi.LIB_PATH, capi._lib)
    capi.LIB_PATH, capi._lib = path, None
    try:
        for name in ('dambreak20_wendland', 'tank30_cubic_dynh', 'tank16_gaussian', 'tank24_wendland_coupled'):
            _parity.test_whole_steps_vs_golden(name)
        _parity.test_dam_break_vs_oracle(150, 'cubic')
        _parity.test_fp32_mode_close_to_fp64()
        _edges.test_cluster_denser_than_the_candidate_list()
        _edges.test_coincident_particles_follow_the_reference_guards()
    finally:
        capi.LIB_PATH, capi._lib = saved


def test_pair_kernel_single_body_variant():
    """-DPAIR_LEAN=0 (the build round 1 was measured with, kept for the A/B run of tools/build_round2_variants.sh): one
    guarded body for every listed pair, the self pair listed.  Same parity bar as the default build."""
    from osph_b200 import capi
    path = emu_build.build(defines=("PAIR_LEAN=0",), tag="_lean0")
    saved = (capi.LIB_PATH, capi._lib)
    capi.LIB_PATH, capi._lib = path, None
    try:
        for name in _parity.STEP_CASES:
            _parity.test_loop_fields_vs_golden(name)
            _parity.test_whole_steps_vs_golden(name)
        _parity.test_fused_loop_equals_explicit_calls('dambreak20_wendland')
        _parity.test_multi_step_call_equals_single_step_calls(None)
        _parity.test_dam_break_vs_oracle(60, 'wendland')
        _parity.test_dam_break_vs_oracle(150, 'cubic')
        _parity.test_dam_break_vs_oracle(150, 'gaussian')
        _parity.test_fp32_mode_close_to_fp64()
        _edges.test_coincident_particles_follow_the_reference_guards()
        _edges.test_cluster_denser_than_the_candidate_list()
        _edges.test_general_lennard_jones_exponents_and_beta_viscosity('cubic')
        _edges.test_domain_far_from_the_origin()
    finally:
        capi.LIB_PATH, capi._lib = saved


@pytest.mark.parametrize("defs,tag", [(("PAIR_UH=0", "PAIR_ISIGN=0", "PAIR_UH_PIPE=0", "PAIR_GEN_PIPE=0"), "_r2gate"),
                                      (("PAIR_UH_PIPE=0", "PAIR_GEN_PIPE=0"), "_uhs"),
                                      (("PAIR_GEN_PIPE=3", "PAIR_PIPE_UNROLL=1", "PAIR_CAP=96"), "_gp3cap96")])
def test_pair_kernel_uniform_h_variant(defs, tag):
    """The default build runs, with Solver(h=value), the pair kernel's uniform-smoothing-length instantiation (loop constants
    instead of the per-pair h terms for fluid neighbours, sign-bit clamps) with a software-pipelined, branch-free flush loop
    (the rare pairs deferred to the end of each flush), and pipelines the general double instantiation as well.  The other
    settings of these knobs: `_r2gate` = the kernel of the round-2 hardware gate before them (no uniform-h instantiation, no
    pipelining, FP64 clamps), `_uhs` = uniform-h without pipelining (bit-identical to the general instantiation), `_gp3cap96`
    = every instantiation pipelined, on the batched staging path.  Same parity bar on every golden case and oracle comparison
    (dynamic h, Gaussian, coupled rows and summation density included), and the bits of the general instantiation
    (OSPH_UH=0) on the same input -- with pipelining up to the summation order of the deferred pairs."""
    from osph_b200 import capi
    path = emu_build.build(defines=defs, tag=tag)
    saved = (capi.LIB_PATH, capi._lib)
    capi.LIB_PATH, capi._lib = path, None
    mp = pytest.MonkeyPatch()
    try:
        for name in _parity.STEP_CASES:
            _parity.test_loop_fields_vs_golden(name)
            _parity.test_whole_steps_vs_golden(name)
        _parity.test_fused_loop_equals_explicit_calls('dambreak20_wendland')
        _parity.test_multi_step_call_equals_single_step_calls(None)
        _parity.test_dam_break_vs_oracle(60, 'wendland')
        _parity.test_dam_break_vs_oracle(150, 'cubic')
        _parity.test_dam_break_vs_oracle(150, 'gaussian')
        _parity.test_fp32_mode_close_to_fp64()
        if "PAIR_UH=0" not in defs:
            for kernel, precision in (('cubic', 'fp64'), ('wendland', 'fp64'), ('cubic', 'fp32'), ('wendland', 'fp32')):
                _parity.test_uniform_h_instantiation_gives_the_bits_of_the_general_one(kernel, precision, mp)
        _edges.test_coincident_particles_follow_the_reference_guards()
        _edges.test_cluster_denser_than_the_candidate_list()
        _edges.test_general_lennard_jones_exponents_and_beta_viscosity('cubic')
        _edges.test_domain_far_from_the_origin()
        _edges.test_fp32_mode_uses_the_reference_cell_rule_too()
    finally:
        mp.undo()
        capi.LIB_PATH, capi._lib = saved


@pytest.mark.parametrize("precision", ['fp64', 'fp32'])
def test_bench_workload_one_step_under_the_guarded_emulator(precision):
    """BASELINE configs[1] itself (dam break N = 1000, 1 009 603 particles: the workload bench.py runs): one whole step of
    the default build under the emulator, whose shared memory and device allocations end at inaccessible pages.  Round 1
    shipped a pair kernel that overran its shared memory only at this size (the wall-row CTAs take the batched staging
    path); this run would have died with SIGSEGV.  The oracle is compared on a strided sample by the GPU test
    tests/test_gpu_fullsize.py; here the step must complete, stay finite and report a clean status."""
    import numpy as np
    from osph_b200 import capi
    from osph_b200 import workloads as W
    case = W.dam_break_case(1000, seed=0)
    cfg = capi.make_config(case['consts'], 'cubic', 'pec', capi.FP64 if precision == 'fp64' else capi.FP32, case['h'])
    with capi.Context(cfg) as ctx:
        ctx.upload(case['pA'])
        ctx.step(1, None, 0.05)
        cols = ctx.download_fields(['ax', 'ay', 'drho', 'rho', 'x'])
        assert ctx.sync() == 0
    fluid = case['pA']['label'] == 0
    for f, c in cols.items():
        assert np.all(np.isfinite(c)), f
    assert np.abs(cols['ax'][fluid]).max() > 0 and np.abs(cols['drho'][fluid]).max() > 0


def test_pair_kernel_packed_scan_variant():
    """-DPAIR_SCAN2=1: the scan phase tests two candidates per packed FP32 instruction (FADD2 / FMUL2 / FFMA2 of sm_100,
    modelled element-wise here) on positions staged as (x0, x1, y0, y1); runs start and end inside a pair, both precisions
    scan on the float copies.  Same parity bar, including the batched staging path (PAIR_CAP=96)."""
    from osph_b200 import capi
    for defs, tag in ((("PAIR_SCAN2=1", "PAIR_LISTPTR=1"), "_scan2"), (("PAIR_SCAN2=1", "PAIR_LISTPTR=1", "PAIR_CAP=96"), "_scan2cap96"),
                      (("PAIR_SCAN2=1", "PAIR_LISTPTR=0", "PAIR_SCAN=8"), "_scan2w8")):
        path = emu_build.build(defines=defs, tag=tag)
        saved = (capi.LIB_PATH, capi._lib)
        capi.LIB_PATH, capi._lib = path, None
        try:
            for name in _parity.STEP_CASES:
                _parity.test_cells_and_neighbour_sets_bit_exact(name)
                _parity.test_whole_steps_vs_golden(name)
            _parity.test_dam_break_vs_oracle(60, 'wendland')
            _parity.test_dam_break_vs_oracle(150, 'cubic')
            _parity.test_dam_break_vs_oracle(150, 'gaussian')
            _parity.test_fp32_mode_close_to_fp64()
            _edges.test_coincident_particles_follow_the_reference_guards()
            _edges.test_cluster_denser_than_the_candidate_list()
            _edges.test_domain_far_from_the_origin()
        finally:
            capi.LIB_PATH, capi._lib = saved


@pytest.mark.parametrize("mode,seed", [(1, 0), (2, 5)])
def test_results_do_not_depend_on_the_schedule(mode, seed):
    """The same parity tests with the emulator resuming the threads of a CTA (and running the CTAs of a launch) in reverse
    and in a seeded random order that changes at every barrier / warp collective: a shared-memory hand-over that is not
    behind a barrier, or a dependence on the order of the grid, shows up as a parity failure here."""
    import ctypes
    from osph_b200 import capi
    lib = ctypes.CDLL(emu_build.build())
    lib.emu_set_order.argtypes = [ctypes.c_int, ctypes.c_ulonglong]
    lib.emu_set_order(mode, seed)
    try:
        _parity.test_cells_and_neighbour_sets_bit_exact('dambreak20_cubic')
        _parity.test_whole_steps_vs_golden('tank30_cubic_dynh')
        _parity.test_whole_steps_vs_golden('tank16_gaussian')
        _parity.test_fused_loop_equals_explicit_calls('dambreak20_wendland')
        _parity.test_dam_break_vs_oracle(60, 'wendland')
        _edges.test_cluster_denser_than_the_candidate_list()
        _solver.test_equation_functions_standalone()
    finally:
        lib.emu_set_order(0, 0)


pytestmark = []          # the star imports brought the modules' `gpu` mark along: these run on the CPU
