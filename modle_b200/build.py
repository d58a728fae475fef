"""Builds libmodle_b200.so (host + CUDA halves of the C ABI) in-tree for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmodle_b200.so")
SOURCES = ["host.cpp", "genome.cpp", "shards.cpp", "kernels.cu", "pixels.cu"]
HEADERS = ["cta.hpp", "sim_types.hpp", "sim_core.hpp", "launch_prep.hpp", "host_rng.hpp",
           "status.hpp", "context.hpp", "ziggurat_tables.inc", os.path.join("..", "..", "include", "modle_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # strict IEEE double arithmetic: no FMA contraction, so results match the CPU oracle
    # (built with -ffp-contract=off) operation for operation
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-cudart", "static", "--threads", "0", "-ldl",
    "-Xlinker", "-soname=libmodle_b200.so",
]


def _deps():
    return [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]


def needs_build():
    from . import buildutil

    return not buildutil.is_current(LIB, [d for d in _deps() if os.path.exists(d)],
                                    " ".join(NVCC_FLAGS))


def build(force=False, verbose=False, variant=None, defines=()):
    """Builds the product library (when its sources changed: content hash, see buildutil), or with
    `variant` an experimental copy libmodle_b200_<variant>.so compiled with extra -D `defines`
    (profiling sessions load it through the MODLE_B200_LIB environment variable)."""
    from . import buildutil

    out = LIB if variant is None else os.path.join(HERE, f"libmodle_b200_{variant}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def cmd(tmp):
        return [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
            ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]

    return buildutil.ensure_built(out, _deps(), cmd, extra=" ".join(NVCC_FLAGS + list(defines)),
                                  force=force or variant is not None, verbose=verbose)


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
