"""
Builds spectralkernels.jl_b200/libsk_b200.so: the C-ABI shared library (include/spectralkernels_b200.h)
with every CUDA kernel compiled for sm_100a.  nvcc cross-compiles without a GPU.

    python spectralkernels.jl_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsk_b200.so")
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("SK_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec); use the system compiler
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
CUDA_HOME = os.path.dirname(os.path.dirname(os.path.realpath(NVCC)))

SOURCES = ["sk_api.cu", "sk_plan_host.cpp"]
HEADERS = ["sk_math.h", "sk_plan.h", "sk_host_util.h", "sk_kernels.cuh", "sk_hankel.h", "sk_hankel.cuh", "sk_rules.cuh", "sk_k8.h", "sk_k8.cuh",
           os.path.join(ROOT, "include", "spectralkernels_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    run = lambda cmd: subprocess.check_call(cmd) if verbose else subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    o_plan = os.path.join(OBJ, "sk_plan_host.o")
    run([GXX, "-O2", "-fPIC", "-fopenmp", "-std=gnu++17", "-fext-numeric-literals", "-ffp-contract=off",
         "-c", os.path.join(CSRC, "sk_plan_host.cpp"), "-o", o_plan])
    o_api = os.path.join(OBJ, "sk_api.o")
    nvflags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-ccbin", GXX, "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr"]
    if verbose:
        nvflags += ["-Xptxas", "-v"]
    run([NVCC, *nvflags, "-c", os.path.join(CSRC, "sk_api.cu"), "-o", o_api])
    run([NVCC, "-shared", "-ccbin", GXX, "-o", LIB, o_api, o_plan,
         "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcufft", "-lquadmath", "-lgomp",
         "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
