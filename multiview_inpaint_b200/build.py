"""Builds the in-tree CUDA library `multiview_inpaint_b200/libgsrast_b200.so` for sm_100a.

Plain nvcc, no torch headers: the library's ABI is the C header include/gsrast_b200.h.
`python -m multiview_inpaint_b200.build [--force] [--verbose]`
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgsrast_b200.so")
SOURCES = ["api.cu", "preprocess.cu", "scan_sort.cu", "binning.cu", "blend_forward.cu",
           "blend_backward.cu", "geom_backward.cu", "geom_backward_multi.cu", "nvls_allreduce.cu", "refstruct.cu", "knn.cu", "loss.cu", "optim.cu", "image_io.cu", "densify.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "--threads", "0"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "gsrast_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgsrast_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
