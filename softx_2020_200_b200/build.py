"""Builds libglsns.so (the C-ABI library of include/glsns.h) in-tree with nvcc for sm_100a.

    python -m softx_2020_200_b200.build [--force]

nvcc cross-compiles without a GPU; the built library is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libglsns.so")
SOURCES = ["api.cu", "assembly.cu", "sparse.cu", "trsv.cu", "krylov.cu", "comm.cu", "host_mesh.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-fopenmp,-O3", "-ccbin", "/usr/bin/g++"] + \
    os.environ.get("GLSNS_NVCC_FLAGS", "").split()   # e.g. -DGLSNS_TRSV_ITEM=96 (tuning experiments)


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".hpp"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "glsns.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "glsns_solver.hpp"))
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(HERE, "_obj", src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", s,
                                                                              "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs +
                              ["-Xcompiler", "-fopenmp", "-ccbin", "/usr/bin/g++", "-lcudart",
                               "-ldl", "-lgomp"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
