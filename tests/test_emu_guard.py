"""CPU: the emulator's shared-memory guard pages.  A kernel that reads past its dynamic or static shared memory must
die under the emulator (round 1: the pair kernel's overrun passed every emulated test and faulted on the hardware)."""
import os
import signal
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
sys.path.insert(0, EMU)
import build as emu_build  # noqa: E402

if not emu_build.available():
    pytest.skip("g++ or the CUDA headers are missing", allow_module_level=True)


@pytest.fixture(scope="module")
def selftest(tmp_path_factory):
    import preprocess
    d = tmp_path_factory.mktemp("emu_guard")
    src = preprocess.transform(open(os.path.join(EMU, "selftest_overrun.cu")).read())
    cpp = d / "selftest_overrun.cpp"
    cpp.write_text('#include "emu.h"\n' + src)
    exe = d / "selftest"
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-Wno-attributes", "-I", EMU, "-I", emu_build.cuda_include(), str(cpp),
                        os.path.join(EMU, "emu.cpp"), "-o", str(exe), "-ldl", "-lrt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def test_in_bounds_kernels_run(selftest):
    r = subprocess.run([selftest, "ok"], capture_output=True, text=True)
    assert r.returncode == 0 and "survived" in r.stdout


@pytest.mark.parametrize("mode", ["dyn", "static"])
def test_shared_memory_overrun_faults(selftest, mode):
    r = subprocess.run([selftest, mode], capture_output=True, text=True)
    assert r.returncode == -signal.SIGSEGV, (r.returncode, r.stdout, r.stderr)
