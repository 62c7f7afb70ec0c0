"""Builds nuradiomc_b200/libnrmc_rt.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libnrmc_rt.so")
SOURCES = [os.path.join(HERE, "csrc", "nrmc_rt.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", f) for f in ("nrmc_math.cuh", "nrmc_att.cuh")] + [os.path.join(REPO, "include", "nrmc_rt.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "--fmad=true"]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc, nvcc wants the system host compiler
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB
