"""In-tree build of the CUDA library (sm_100a only) with nvcc.

`python -m chad_tsdf_b200.build` or `build_library()`; the resulting chad_tsdf_b200/libchad_b200.so
is git-ignored but travels with the tree. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libchad_b200.so")
SOURCES = ["radix_sort.cu", "points.cu", "shard.cu", "band.cu", "blocks.cu", "runs.cu", "fold.cu", "dag.cu", "context.cu", "nccl_dyn.cpp", "tsdf_host.cpp"]
HEADERS = ["common.cuh", "kernels.cuh", "points.cuh", "radix_sort.cuh", "scan.cuh", "ray.cuh", "nccl_dyn.h"]
# -fmad=false / -ffp-contract=off: the reference's strict-IEEE configuration (cmake/options_compiler.cmake:39);
# the kernels additionally use explicit *_rn intrinsics wherever a result is observable.
NVCC_FLAGS = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-cudart", "static"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "chad_b200.h"), os.path.join(ROOT, "include", "chad", "tsdf.hpp")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_facade_demo(verbose: bool = False) -> str:
    """tests/cpp/facade_demo.cpp against include/chad/tsdf.hpp + libchad_b200.so (the drop-in check)."""
    out = os.path.join(ROOT, "tests", "cpp", "facade_demo")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"),
           "-o", out, "-L", PKG, "-lchad_b200", f"-Wl,-rpath,{PKG}"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def build_facade_overloads(verbose: bool = False) -> str:
    """tests/cpp/facade_overloads.cpp: every insert() overload / constructor of the reference's class, the glm and Eigen ones
    compiled against the stand-in headers under tests/cpp/shims (neither library is installed here)."""
    out = os.path.join(ROOT, "tests", "cpp", "facade_overloads")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++20", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp", "shims"),
           os.path.join(ROOT, "tests", "cpp", "facade_overloads.cpp"), "-o", out, "-L", PKG, "-lchad_b200", f"-Wl,-rpath,{PKG}"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def build_facade_persist(verbose: bool = False) -> str:
    """tests/cpp/facade_persist.cpp: save -> load -> continue, the leaf iterator (host cursor and device), page-locked vectors."""
    out = os.path.join(ROOT, "tests", "cpp", "facade_persist")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++20", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp", "shims"),
           os.path.join(ROOT, "tests", "cpp", "facade_persist.cpp"), "-o", out, "-L", PKG, "-lchad_b200", f"-Wl,-rpath,{PKG}"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def build_dag_reader(verbose: bool = False) -> str:
    """tests/cpp/dag_reader.cpp: chad::load_dag + HostNodeLevels::query (host-only code of the library; runs without a GPU)."""
    out = os.path.join(ROOT, "tests", "cpp", "dag_reader")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++20", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dag_reader.cpp"),
           "-o", out, "-L", PKG, "-lchad_b200", f"-Wl,-rpath,{PKG}"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
