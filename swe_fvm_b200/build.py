"""In-tree build of libswe_b200.so (nvcc, sm_100a only). The .so is git-ignored but travels
to the GPU box with the working tree."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libswe_b200.so")
SOURCES = ["swe_b200.cu", "hostmesh.cpp", "distplan.cpp"]
HEADERS = ["swe_kernels.cuh", "swe_device.cuh", "swe_cases.cuh", "swe_dist.cuh", "swe_flux_registry.cuh", "user_fluxes.cuh", "hostmesh.hpp", "distplan.hpp", os.path.join("..", "..", "include", "swe_b200.h"), os.path.join("..", "..", "include", "swe_constants.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # bit parity with the CPU oracle (g++ -ffp-contract=off)
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2,-fopenmp",
    "-shared", "-cudart", "static", "-lgomp",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_variant(out: str, defines: list[str], verbose: bool = False) -> str:
    """Tuning builds: same sources with extra -D flags, written to `out`."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if os.environ.get("SWE_B200_LIB"):
        return os.environ["SWE_B200_LIB"]
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return LIB


def pybind_path() -> str:
    import sysconfig
    return os.path.join(os.path.dirname(_HERE), "pybind", "SWE_FVM" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pybind(force: bool = False) -> str:
    """The pybind11 module SWE_FVM (pybind/swe_fvm_module.cpp) over the C++ host API; host-only C++."""
    import sysconfig
    import pybind11
    root = os.path.dirname(_HERE)
    src = os.path.join(root, "pybind", "swe_fvm_module.cpp")
    out = pybind_path()
    deps = [src, LIB] + [os.path.join(root, "include", "swe", f) for f in os.listdir(os.path.join(root, "include", "swe"))]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I" + pybind11.get_include(),
           "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(root, "include"), src, "-L" + _HERE, "-lswe_b200",
           "-Wl,-rpath," + _HERE, "-o", out]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    build_lib(force=True, verbose="-v" in sys.argv)
    print(LIB)
