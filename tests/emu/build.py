"""Builds tests/emu/libosph_emu.so: the kernel sources of offshore-sph_b200/csrc compiled by g++ against the SIMT
emulator (emu.h / emu.cpp) instead of nvcc + libcudart.  TEST INFRASTRUCTURE ONLY: the product never loads this file
(osph_b200.capi looks for lib/libosph_b200.so); tests/test_emu_*.py point the binding at it explicitly.

    python tests/emu/build.py [-f]
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "offshore-sph_b200", "csrc")
GEN = os.path.join(HERE, "gen")
LIB = os.path.join(HERE, "libosph_emu.so")
sys.path.insert(0, HERE)


def cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    nvcc = shutil.which("nvcc")
    if nvcc:
        return os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
    return None


def available():
    return shutil.which("g++") is not None and cuda_include() is not None


def build(force=False, verbose=False, defines=(), tag=""):
    """defines/tag: a variant build (e.g. a tiny PAIR_CAP to force the pair kernel's batched staging path) -> libosph_emu<tag>.so"""
    import preprocess
    LIB = os.path.join(HERE, "libosph_emu%s.so" % tag)
    GEN = os.path.join(HERE, "gen" + tag)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(HERE, f) for f in ("emu.h", "emu.cpp", "preprocess.py", "build.py")]
    srcs.append(os.path.join(ROOT, "include", "osph.h"))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in srcs):
        return LIB
    gen = preprocess.main(CSRC, GEN)
    units = [g for g in gen if g.endswith(".cpp")] + [os.path.join(HERE, "emu.cpp")]
    flags = ["-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas",
             "-Wno-attributes", "-Wno-deprecated-declarations", "-I", HERE, "-I", CSRC, "-I", cuda_include()]
    flags += ["-D" + d for d in defines]
    objs = []

    def compile_one(u):
        o = os.path.join(GEN, os.path.basename(u) + ".o")
        r = subprocess.run(["g++"] + flags + ["-c", u, "-o", o], capture_output=True, text=True)
        return u, o, r

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for u, o, r in ex.map(compile_one, units):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("g++ failed on " + u)
            objs.append(o)
    r = subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-Wl,--no-undefined", "-o", LIB] + objs + ["-ldl", "-lrt"],
                       capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("link of libosph_emu.so failed")
    return LIB


def build_fake_nccl():
    """tests/emu/libfake_nccl.so: the NCCL calls of slab_nccl.cu over shared memory between rank processes (CPU tests)."""
    src, out = os.path.join(HERE, "fake_nccl.cpp"), os.path.join(HERE, "libfake_nccl.so")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I", cuda_include(), src, "-o", out, "-lrt"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed on fake_nccl.cpp")
    return out


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
