#!/usr/bin/env python
"""Benchmark of the B200-native WCSPH step (BASELINE.json metric: particle-steps/s, 2-D dam break).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp64|fp32]
                    [--particles-per-side N] [--kernel cubic|wendland|gaussian]

One JSON line on stdout (rank 0).  A "step" is one whole WCSPH time step (dt reduction, PEC predict,
neighbour structure, fused pair kernel, PEC correct) over the synthetic dam-break particle block of
SURVEY.md section 8(d): the reference's DamBreak generator scaled to N x N fluid particles, setup as
Solver.setup() does, jittered deterministically so pair forces do not vanish.

  value          particle-steps/s with the state resident in HBM: ONE osph_step(K) call for the K timed steps, no host
                 round trip inside (the library applies the corrector of step k in the predictor pass of step k+1)
  e2e            the same through the plugin boundary with HOST buffers: every step uploads the packed
                 particle_dtype array from pinned memory, runs one step, downloads it again
  roofline       the fused pair kernel against the measured HBM peak (algorithmic bytes 13F+1 per particle)
  cpu_baseline   the CPU oracle (a C port of the reference's numba path, pair loop on all host threads): a full step where it
                 fits the budget (1 M particles), else a strided sample scaled back -- the line says which
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "offshore-sph_b200"))

METRIC = "particle-steps/s, 2D WCSPH dam break"
UNIT = "particle-steps/s"
DAMPING = 0.05


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


WORKLOAD = "dam_break"      # set from --workload: dam_break (BASELINE configs[0,1,4]) | containment (configs[2]) | icebreak (configs[3])


def build_case(n_side, seed=0):
    """Synthetic particle block after Solver.setup() + deterministic jitter."""
    from osph_b200 import workloads as W
    if WORKLOAD == "containment":          # closed tank, dynamic h = 1.3 sqrt(m/rho), XSPH off (examples/Containment.py)
        return W.tank_case(n_side, h=None, useXSPH=False, seed=seed)
    if WORKLOAD == "icebreak":             # tank with a floating row of Coupled particles (examples/IceBreak.py geometry)
        return W.tank_case(n_side, width=60.0, height=10.0, h=1.3 * 60.0 / n_side, useXSPH=True, seed=seed, coupled_row=True)
    return W.dam_break_case(n_side, seed=seed)


def resolve_problem(args, world):
    """(particles per side of the generator, "weak" | "strong").  One GPU: BASELINE configs[1], N = 1000 (1 M particles).
    Several GPUs: the FIXED 16 M-particle dam break of configs[4] (N = 4000), i.e. strong scaling of the north star's
    problem, unless --weak asks for 1 M particles per GPU (N = 1000 sqrt(gpus))."""
    if world <= 1:
        return (args.particles_per_side or 1000), "weak"
    if args.weak:
        return int(round((args.particles_per_side or 1000) * (world ** 0.5))), "weak"
    return (args.particles_per_side or 4000), "strong"


def workload_name(n_side, n, kernel, prec):
    k = {'cubic': 'cubic spline', 'wendland': 'Wendland', 'gaussian': 'Gaussian'}[kernel]
    if WORKLOAD == "containment":
        return "containment tank Nx=%d (%d particles), %s, PEC, no XSPH, dynamic h, %s" % (n_side, n, k, prec)
    if WORKLOAD == "icebreak":
        return "ice-break tank Nx=%d (%d particles incl. coupled ice row), %s, PEC, XSPH, %s" % (n_side, n, k, prec)
    return "dam_break N=%d (%d particles), %s, PEC, XSPH, h=1.6 r0, %s" % (n_side, n, k, prec)


# ------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def host_threads():
    """Host threads the CPU legs may use: the cores this process is allowed on."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def time_oracle_step(case, kernel, budget_s=12.0, threads=None):
    """One step of the oracle port with its pair loop on `threads` host threads (default: all the process may use; the
    reference itself is single-threaded numba, the other phases run on one thread and are 0.1 % of the step).  When a
    full step does not fit the budget the pair loop runs on every `stride`-th fluid particle and is scaled back.
    Returns (seconds per full step, sample text)."""
    from oracle import oracle as O
    threads = threads or host_threads()
    pA, c = case['pA'], case['consts']
    n_f = int((pA['label'] == 0).sum())
    # cost model of the reference search: 9 cells of 1 m^2 -> candidates per particle (one thread: 3-4 ns per candidate,
    # 36 ns per accepted pair, measured with this port at 1 M and 16 M; threads scale it by about 0.7 per thread)
    cand = 9.0 * max(1.0, n_f / 625.0)
    est = (n_f * cand * 4e-9 + n_f * 72 * 36e-9) / max(1.0, 0.7 * threads)
    stride = max(1, int(np.ceil(est / budget_s)))
    time_oracle_step.last_stride = stride
    time_oracle_step.last_threads = threads
    cache = getattr(time_oracle_step, "_cache", None)
    if cache is not None and cache[0] is case and cache[1] == kernel:
        # a later step of the same run: the phases outside the pair loop (time step, predictor, grid, corrector: 0.1 % of a
        # step, all particles, one thread) are not repeated -- their time from the first step is added -- so that a run of
        # many steps on 16 M particles stays within minutes; the pair loop is timed again on the same predicted state
        _, _, P, w, grid, t_other = cache
        t1 = time.perf_counter()
        pairs = O.loop(P, w, grid, kernel, stride, 0, threads=threads)
        t2 = time.perf_counter()
        other_txt = "%.2f s, timed at the first step" % t_other
    else:
        P = O.Particles.from_aos(pA)
        w = O.wcsph(c['height'], c['r0'], c['rho0'], True)
        fluid = P.fluid
        t0 = time.perf_counter()
        dt3 = O.timestep(P, fluid)
        O.pec_predict(P, fluid, dt3[0], DAMPING, True, False)
        grid = O.Grid(P, 2.0)
        P.h[fluid.astype(bool)] = case['h']
        t1 = time.perf_counter()
        pairs = O.loop(P, w, grid, kernel, stride, 0, threads=threads)
        t2 = time.perf_counter()
        # the corrector on a copy of the integrated fields: P stays the predicted state for the steps that follow
        keep = {f: getattr(P, f).copy() for f in ('x', 'y', 'vx', 'vy', 'rho')}
        O.pec_correct(P, fluid, dt3[0], DAMPING, True, False)
        t3 = time.perf_counter()
        for f, v in keep.items():
            getattr(P, f)[:] = v
        t_other = (t1 - t0) + (t3 - t2)
        time_oracle_step._cache = (case, kernel, P, w, grid, t_other)
        other_txt = "%.2f s" % t_other
    t_full = t_other + (t2 - t1) * stride
    if stride == 1:
        sample = ("1 full step on all %d particles: pair loop on %d host threads (%d pairs, %.1f s), other phases on one "
                  "thread (%s)" % (P.n, threads, pairs, t2 - t1, other_txt))
    else:
        sample = ("1 step; pair loop on %d host threads over every %d-th fluid particle (%d of %d, %d pairs, %.1f s) scaled "
                  "by %d; other phases on all %d particles (%s)" % (threads, stride, (n_f + stride - 1) // stride, n_f,
                                                                    pairs, t2 - t1, stride, P.n, other_txt))
    return t_full, sample


def cpu_baseline_entry(value, sample):
    """The cpu_baseline object of a bench line, from the last time_oracle_step call."""
    stride = time_oracle_step.last_stride
    return {"value": value, "unit": UNIT, "cores": time_oracle_step.last_threads, "kind": "port", "sample": sample,
            "extrapolated": stride > 1, "stride": stride, "host_cores_available": os.cpu_count(),
            "note": "C port of the reference's numba loop (the reference is single-threaded and about 5x slower per "
                    "thread, BASELINE.md); the pair loop is spread over the host threads"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; pair loop on all host threads the process may
    use -- the reference's own numba loop is single-threaded) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same workload as the GPU arm at this GPU count
    n_side, scaling = resolve_problem(args, args.gpus)
    case = build_case(n_side)
    n = len(case['pA'])
    from oracle import oracle as O
    O.build()
    times = []
    sample = ""
    budget = max(0.5, min(40.0, 180.0 / max(1, args.steps + args.warmup)))      # the whole run stays near 3 minutes of CPU work
    for s in range(args.warmup + args.steps):
        t, sample = time_oracle_step(case, args.kernel, budget)
        if s >= args.warmup:
            times.append(t)
    t_step = float(np.mean(times))
    value = n / t_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n_side, n, args.kernel, "FP64"), "particles": n},
        # stride > 1: the denominator is a MODEL of a full step, not a measured one -- the pair loop ran on every stride-th
        # fluid particle and its time was multiplied by the stride (a full step is hours at 16 M); stride 1: measured
        "extrapolated": time_oracle_step.last_stride > 1, "stride": time_oracle_step.last_stride,
        "cpu_baseline": cpu_baseline_entry(value, sample),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling in the background.  Started BEFORE the warm-up (the tool needs a few hundred ms to emit its
    first line); samples are time-stamped and only those inside [mark_start, mark_end] -- the timed region -- count."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.5)
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.t1 = time.time()
        time.sleep(0.1)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                parsed.append((ts, float(r[2]), float(r[3]), [nm for k, nm in enumerate(names)
                                                             if r[6 + k].strip().lower().startswith("active")]))
            except Exception:
                pass
        inside = [q for q in parsed if self.t0 is not None and self.t0 - 0.02 <= q[0] <= self.t1 + 0.02]
        use = inside if inside else parsed[-3:]
        sm = [q[1] for q in use]; mx = [q[2] for q in use]
        reasons = sorted({nm for q in use for nm in q[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(inside), "samples_total": len(parsed),
                "window_s": (self.t1 - self.t0) if self.t0 is not None else None}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from osph_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the WCSPH step has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        from osph_b200 import slabs

        def cpu_leg(case):
            if args.no_cpu_baseline:
                return None
            from oracle import oracle as O
            O.build()
            t_cpu, sample = time_oracle_step(case, args.kernel, args.cpu_budget)
            return cpu_baseline_entry(len(case['pA']) / t_cpu, sample)
        return slabs.bench_multi_gpu(args, rank, world, local, cpu_baseline=cpu_leg)

    prec = capi.FP64 if args.precision == "fp64" else capi.FP32
    F = 8 if prec == capi.FP64 else 4
    n_side, _ = resolve_problem(args, 1)
    case = build_case(n_side)
    pA, c = case['pA'], case['consts']
    n = len(pA)
    cfg = capi.make_config(c, args.kernel, 'pec', prec, case['h'], device=local)
    ctx = capi.Context(cfg)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)

    def timed_region(fn, iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    # ---- resident-in-HBM throughput -----------------------------------------------------------
    clocks = ClockSampler(local)
    ctx.upload(pA)
    ctx.step(args.warmup, None, DAMPING)
    ctx.sync(); ctx.pair_kernel_time()
    l0 = ctx.launch_count
    builds0, sorts0 = ctx.sort_stats()
    clocks.mark_start()
    # one osph_step call for all K steps: the library fuses the corrector of step k with the predictor of step k+1
    t_dev = timed_region(lambda: ctx.step(args.steps, None, DAMPING), 1)
    clocks.mark_end()
    clk = clocks.stop()
    launches = ctx.launch_count - l0
    pair_us, pair_n = ctx.pair_kernel_time()
    status = ctx.sync()
    builds1, sorts1 = ctx.sort_stats()
    pk_launches, pk_uniform, pk_flags = ctx.pair_kernel_info()
    value = n * args.steps / t_dev

    # ---- end to end through the plugin boundary, host buffers ---------------------------------
    hbuf = torch.empty(n * 154, dtype=torch.uint8, pin_memory=True)
    host = hbuf.numpy().view(pA.dtype)
    host[:] = ctx.download(pA.copy())
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        ctx.upload(host)
        ctx.step(1, None, DAMPING)
        ctx.download(host)

    if args.no_e2e:                      # profiler / sanitizer runs only: a line without e2e is not a bench line
        e2e_steps, e2e_value = 0, None
    else:
        e2e_step()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        t_e2e = time.perf_counter() - t0
        e2e_value = n * e2e_steps / t_e2e

    # ---- roofline of the dominant kernel (fused pair kernel) -----------------------------------
    peaks, which = measured_peaks()
    alg_bytes = (13 * F + 1) * n
    achieved = alg_bytes / (pair_us * 1e-6) / 1e9 if pair_us > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")) as f:
            t = json.load(f).get("%s_%s_N%d" % (args.precision, args.kernel, n_side))
            traffic = t["dram_bytes_per_launch"] if t else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_pair (fused EOS-input/continuity/momentum/viscosity/XSPH/LJ)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "avg_launch_us": pair_us, "launches_timed": pair_n, "share_of_step": pair_us * 1e-6 * args.steps / t_dev,
                "algorithmic_bytes_per_particle": 13 * F + 1,
                "note": "k_pair is bound by the FP pipe / dependent-instruction latency, not by HBM (DESIGN.md section 4): "
                        "fp_pipe is the fraction to read first, frac (HBM, as the contract defines it) second"}
    # FP-pipe view of the same kernel: SURVEY 8(d) flop model (70 flop per contributing pair, 10 per rejected
    # candidate) against the FMA peak measured with tools/fp64_pipe.cu on this pool
    try:
        fp = json.load(open(os.path.join(ROOT, "profiles", "measured_fp_peaks.json")))
        q = 3.0 if args.kernel == "gaussian" else 2.0
        pairs = np.pi * (q * 1.6) ** 2                       # contributing neighbours per fluid particle at h = 1.6 r0
        cand = 9.0 * (q * 1.6) ** 2                          # 3x3 cells of the pair radius
        flops = (70.0 * pairs + 10.0 * (cand - pairs)) * int((pA['label'] == 0).sum())
        peak = fp["fp64_tflops"] if prec == capi.FP64 else fp["fp32_tflops"]
        ach = flops / (pair_us * 1e-6) / 1e12 if pair_us > 0 else 0.0
        roofline["fp64_tflops_peak"] = fp["fp64_tflops"]; roofline["fp32_tflops_peak"] = fp["fp32_tflops"]
        roofline["primary"] = "fp_pipe"
        roofline["fp_pipe"] = {"achieved_tflops_model": ach, "peak_tflops_measured": peak, "frac": ach / peak,
                               "pair_interactions_per_s": pairs * int((pA['label'] == 0).sum()) / (pair_us * 1e-6) if pair_us > 0 else 0.0,
                               "model": "70 flop x %.1f contributing pairs + 10 flop x %.1f rejected candidates per fluid particle" % (pairs, cand - pairs)}
    except Exception:
        pass
    # ---- CPU baseline: oracle port on a bounded sample ----------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        t_cpu, sample = time_oracle_step(case, args.kernel, args.cpu_budget)
        cpu = cpu_baseline_entry(n / t_cpu, sample)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if prec == capi.FP64 else "f32 pair arithmetic, f64 state", "data": "synthetic",
        "config": {"workload": workload_name(n_side, n, args.kernel, args.precision.upper()),
                   "particles": n, "fluid": int((pA['label'] == 0).sum()), "damping": DAMPING, "dt": "dynamic",
                   "l2": "per-step working set (%.0f MB state + sorted copies) exceeds the 126 MB L2" % (n * 19 * 8 / 1e6),
                   "neighbour_structure": "counting sort by cell; %d of the %d timed builds sorted, the others reused the binning "
                                          "(cells = pair radius + skin, OSPH_SKIN)" % (sorts1 - sorts0, builds1 - builds0),
                   "pair_kernel": ("uniform-smoothing-length instantiation (Solver(h=value): h_ij terms of a fluid-fluid pair are loop "
                                   "constants)" if pk_uniform == pk_launches else "general instantiation") +
                                  (", software-pipelined flush loop" if pk_flags & (4 if pk_uniform == pk_launches else 8) else ""),
                   "parallelism": "1 GPU"},
        "clocks": clk, "gpu_launches": launches, "status_bits": status,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 154, "d2h_bytes_per_step": n * 154,
                "steps": e2e_steps, "what": "osph_upload_aos(pinned) + osph_step(1) + osph_download_aos(pinned), host wall clock"},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    ctx.close()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--kernel", default="cubic", choices=["cubic", "wendland", "gaussian"])
    ap.add_argument("--particles-per-side", type=int, default=None,
                    help="N of the dam-break generator (N x N fluid particles).  Default: 1000 on one GPU (BASELINE configs[1]); "
                         "4000 = the fixed 16 M-particle problem of configs[4] on several GPUs; with --weak it is per GPU")
    ap.add_argument("--workload", default="dam_break", choices=["dam_break", "containment", "icebreak"],
                    help="dam_break = BASELINE configs[0,1,4] (default, the metric's workload); containment = configs[2] "
                         "(use --particles-per-side 2000 --precision fp32); icebreak = configs[3] (--particles-per-side 4900)")
    ap.add_argument("--weak", action="store_true",
                    help="multi-GPU: weak scaling, --particles-per-side (default 1000) is per GPU, N*sqrt(gpus) in total; "
                         "the default on several GPUs is strong scaling of the fixed 16 M-particle problem")
    ap.add_argument("--total-side", action="store_true", help="accepted for compatibility: strong scaling is the default now")
    ap.add_argument("--sequencer", default="p2p", choices=["p2p", "nccl", "python"],
                    help="multi-GPU exchange: p2p = NVLink peer-memory windows + mailbox kernels (default, falls back to "
                         "nccl), nccl = direct NCCL calls inside the library, python = osph_b200/slabs.py via torch.distributed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiler and sanitizer runs)")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
