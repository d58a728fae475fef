"""Builds tests/cabi/consumer (a C++ program that includes include/modle_b200.h and links
libmodle_b200.so). TEST INFRASTRUCTURE."""
import os

from modle_b200 import build, buildutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cabi", "consumer.cpp")
OUT = os.path.join(HERE, "cabi", "consumer")
CXXFLAGS = ["-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror"]


def _build(src, out, extra_flags, force):
    lib = build.build()
    cxx = os.environ.get("CXX", "g++")
    deps = [src, os.path.join(ROOT, "include", "modle_b200.h")]

    def cmd(tmp):
        # linked by soname + an $ORIGIN-relative rpath: the tree is copied to the GPU box
        return [cxx] + CXXFLAGS + ["-I", os.path.join(ROOT, "include"), src, "-o", tmp,
                                   "-L", os.path.dirname(lib), "-lmodle_b200",
                                   "-Wl,-rpath,$ORIGIN/../../modle_b200"] + extra_flags

    return buildutil.ensure_built(out, deps, cmd, extra=" ".join(CXXFLAGS + extra_flags),
                                  force=force)


def build_consumer(force=False):
    return _build(SRC, OUT, [], force)


def build_consumer_multi(force=False):
    """The multi-GPU consumer needs the CUDA runtime and NCCL (system libnccl, /usr/include/nccl.h)
    like any C++ host that drives several GPUs would."""
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return _build(os.path.join(HERE, "cabi", "consumer_multi.cpp"),
                  os.path.join(HERE, "cabi", "consumer_multi"),
                  ["-I", os.path.join(cuda, "include"), "-L", os.path.join(cuda, "lib64"),
                   "-lcudart", "-lnccl", "-Wl,-rpath," + os.path.join(cuda, "lib64")], force)
