"""Build recipe for libcgvec_b200.so (in-tree, sm_100a only) and the C++ host-mirror demo.

nvcc cross-compiles here without a GPU; the built .so is git-ignored but travels to the GPU box with
the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libcgvec_b200.so")
DEMO = os.path.join(PKG, "host", "cgvec_host_demo")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def lib_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".inl"))]
    srcs.append(os.path.join(ROOT, "include", "cgvec.h"))
    return srcs


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = lib_sources()
    if not force and _newer(LIB, srcs):
        return LIB
    cu = [s for s in srcs if s.endswith(".cu")]
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *cu, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)       # the image's CC wrapper lacks pieces nvcc's host pass needs
    subprocess.run(cmd, check=True, env=env)
    return LIB


def build_host_demo(force: bool = False) -> str:
    """g++ build of the C++ host mirror (VectorStore / SurrealVectorBackend / SemanticSearch over the C ABI)."""
    src = os.path.join(PKG, "host", "demo_main.cpp")
    hdrs = [os.path.join(PKG, "host", f) for f in os.listdir(os.path.join(PKG, "host")) if f.endswith(".hpp")]
    if not os.path.exists(src):
        return ""
    if not force and _newer(DEMO, [src, *hdrs, LIB]):
        return DEMO
    cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"), src,
           "-o", DEMO, "-L", PKG, "-lcgvec_b200", "-Wl,-rpath,$ORIGIN/..", "-ldl", "-lpthread"]
    subprocess.run(cmd, check=True)
    return DEMO


MOCK_TEST = os.path.join(PKG, "host", "mock_backend_test")


def build_mock_test(force: bool = False) -> str:
    """CPU-only replay of the reference's MockBackend unit tests against the C++ host mirror (links the C ABI library
    only because the header references its symbols; it never creates an index)."""
    src = os.path.join(PKG, "host", "mock_backend_test.cpp")
    hdr = os.path.join(PKG, "host", "cgvec_host.hpp")
    if not force and _newer(MOCK_TEST, [src, hdr, LIB]):
        return MOCK_TEST
    cmd = ["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"), src, "-o", MOCK_TEST,
           "-L", PKG, "-lcgvec_b200", "-Wl,-rpath,$ORIGIN/..", "-ldl", "-lpthread"]
    subprocess.run(cmd, check=True)
    return MOCK_TEST


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_host_demo(force)
    build_mock_test(force)
