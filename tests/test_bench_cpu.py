"""CPU: the legs of bench.py that need no GPU -- the reference arm (`--impl reference`: the oracle port on the host cores)
and the line's contract keys; the GPU arm must refuse to run without a device (there is no CPU path)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--particles-per-side", "60", "--steps", "3", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("particle-steps/s") and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - d["config"]["particles"]) < 1e-6 * d["config"]["particles"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    # a step of 3 843 particles fits any budget: measured in full, not extrapolated
    assert cb["stride"] == 1 and cb["extrapolated"] is False and d["extrapolated"] is False
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_runs_on_rank_zero_only():
    out = _run("--impl", "reference", "--gpus", "2", "--particles-per-side", "40", "--steps", "1", "--warmup", "0",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_later_steps_of_a_run_re_time_the_pair_loop_only():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    O.build()
    case = bench.build_case(40)
    t1, s1 = bench.time_oracle_step(case, "cubic", 5.0, threads=2)
    t2, s2 = bench.time_oracle_step(case, "cubic", 5.0, threads=2)
    assert "timed at the first step" in s2 and "timed at the first step" not in s1
    assert t1 > 0 and t2 > 0 and bench.time_oracle_step.last_threads == 2
    # another case is a new run
    _, s3 = bench.time_oracle_step(bench.build_case(30), "cubic", 5.0, threads=1)
    assert "timed at the first step" not in s3


def test_gpu_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        return
    out = _run("--steps", "1", "--warmup", "1")
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
