"""Builds libitr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python image-text-retrieval_b200/build.py [--force] [--verbose]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libitr_b200.so")
SOURCES = ["simt_kernels.cu", "vse_step.cu", "scan_bwd.cu", "scan_t2i_tc.cu", "scan_t2i_tc2.cu", "scan_i2t_tc2.cu", "tc_microbench2.cu", "plan.cpp"]
HEADERS = ["common.cuh", "scan_f32.cuh", "tc_ptx.cuh", "tc2_common.cuh", os.path.join("..", "..", "include", "itr_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Every source is compiled to its own object (in parallel), then linked into the shared library."""
    if not force and not is_stale():
        return LIB
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    with tempfile.TemporaryDirectory(prefix="itr_b200_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, os.path.splitext(src)[0] + ".o")
            cmd = [_nvcc()] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            return obj, subprocess.run(cmd, capture_output=True, text=True)
        with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
            results = list(pool.map(compile_one, SOURCES))
        for obj, res in results:
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed compiling {} (exit {})".format(os.path.basename(obj), res.returncode))
        link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC"] + \
               [obj for obj, _ in results] + ["-o", LIB]
        res = subprocess.run(link, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed linking libitr_b200.so (exit {})".format(res.returncode))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
