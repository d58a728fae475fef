"""Builds tests/cabi/consumer (a C++ program that includes include/modle_b200.h and links
libmodle_b200.so). TEST INFRASTRUCTURE."""
import os

from modle_b200 import build, buildutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cabi", "consumer.cpp")
OUT = os.path.join(HERE, "cabi", "consumer")
CXXFLAGS = ["-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror"]


def build_consumer(force=False):
    lib = build.build()
    cxx = os.environ.get("CXX", "g++")
    deps = [SRC, os.path.join(ROOT, "include", "modle_b200.h")]

    def cmd(tmp):
        # $ORIGIN-relative rpath: the tree is copied to the GPU box
        return [cxx] + CXXFLAGS + ["-I", os.path.join(ROOT, "include"), SRC, "-o", tmp, lib,
                                   "-Wl,-rpath,$ORIGIN/../../modle_b200"]

    return buildutil.ensure_built(OUT, deps, cmd, extra=" ".join(CXXFLAGS), force=force)
