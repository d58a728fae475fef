"""Builds libmodle_b200.so (host + CUDA halves of the C ABI) in-tree for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmodle_b200.so")
SOURCES = ["host.cpp", "genome.cpp", "kernels.cu", "pixels.cu"]
HEADERS = ["cta.hpp", "sim_types.hpp", "sim_core.hpp", "launch_prep.hpp", "host_rng.hpp",
           "status.hpp", "context.hpp", os.path.join("..", "..", "include", "modle_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # strict IEEE double arithmetic: no FMA contraction, so results match the CPU oracle
    # (built with -ffp-contract=off) operation for operation
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-cudart", "static", "--threads", "0",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, variant=None, defines=()):
    """Builds the product library, or with `variant` an experimental copy
    libmodle_b200_<variant>.so compiled with extra -D `defines` (profiling sessions load it
    through the MODLE_B200_LIB environment variable)."""
    out = LIB if variant is None else os.path.join(HERE, f"libmodle_b200_{variant}.so")
    if variant is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building " + out)
    return out


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
