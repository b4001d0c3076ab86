"""Builds tests/mirror/libglsns_mirror.so: the transcribed reference drivers (reference_drivers.hpp)
around the product's host-side solver class, linked against the product library.  Test harness."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libglsns_mirror.so")
PRODUCT_DIR = os.path.join(ROOT, "softx_2020_200_b200")


def build(force=False):
    src = os.path.join(HERE, "host_solver.cpp")
    deps = [src, os.path.join(HERE, "reference_drivers.hpp"),
            os.path.join(ROOT, "include", "glsns_solver.hpp"), os.path.join(ROOT, "include", "glsns.h"),
            os.path.join(PRODUCT_DIR, "libglsns.so")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB)
                                                 for d in deps if os.path.exists(d)):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", LIB,
                           "-L" + PRODUCT_DIR, "-l:libglsns.so",
                           "-Wl,-rpath," + PRODUCT_DIR, "-Wl,-rpath,$ORIGIN/../../softx_2020_200_b200"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
