"""In-tree build of the CUDA library (sm_100a) and staging of input data.

    python -m rawhash_b200.build            # builds rawhash_b200/librawhash_b200.so

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librawhash_b200.so")
CLI = os.path.join(HERE, "rawhash2_b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # no implicit FMA: every fused op in the kernels is an explicit __fmaf_rn/__fma_rn
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
]
COMPILE_FLAGS = NVCC_FLAGS
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared"]
SOURCES = ["rh_gpu.cu", "rh_index_gpu.cu", "rh_host.cpp", "rh_io.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join(ROOT, "include", "rawhash_b200.h")]


def _newer(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """One object per source under csrc/_obj/ (rebuilt when the source or any header is newer), then one link."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    obj_dir = os.path.join(CSRC, "_obj")
    os.makedirs(obj_dir, exist_ok=True)
    objs, relink = [], force or not os.path.isfile(LIB)
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(obj_dir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            _run([nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj], verbose)
            relink = True
    if relink or _newer(LIB, objs):
        _run([nvcc] + LINK_FLAGS + ["-o", LIB] + objs + ["-lz", "-ldl"], verbose)
    return LIB


def build_cli(force: bool = False) -> str:
    """rawhash_b200/rawhash2_b200: the `rawhash2` command line (src/main.cpp) on top of the C-ABI library."""
    src = os.path.join(CSRC, "rh_main.cpp")
    if force or _newer(CLI, [src, LIB, os.path.join(ROOT, "include", "rawhash_b200.h")]):
        gxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
        _run([gxx, "-std=c++17", "-O2", "-pthread", "-Wall", "-o", CLI, src, "-L" + HERE, "-lrawhash_b200", "-Wl,-rpath,$ORIGIN"], False)
    return CLI


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:1] + cmd[-3:]))
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)


def stage_models() -> None:
    """Copy the ONT k-mer model tables the reference vendors into data/models/ (git-ignored)."""
    srcs = {
        "r9.4_6mer.model": "/root/reference/extern/kmer_models/legacy/legacy_r9.4_180mv_450bps_6mer/template_median68pA.model",
        "r10.4.1_9mer.txt": "/root/reference/extern/kmer_models/dna_r10.4.1_e8.2_400bps/9mer_levels_v1.txt",
    }
    dst_dir = os.path.join(ROOT, "data", "models")
    os.makedirs(dst_dir, exist_ok=True)
    for name, src in srcs.items():
        dst = os.path.join(dst_dir, name)
        if os.path.isfile(src) and not os.path.isfile(dst):
            shutil.copyfile(src, dst)


def build_oracle() -> None:
    """Build the checkers (oracle/librh_oracle.so and, when the reference tree exists, oracle/_ref)."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("oracle build failed")


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_cli(force="--force" in sys.argv)
    stage_models()
    print(LIB)
