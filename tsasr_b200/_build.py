"""Compiles libtsasr_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The library exposes a plain C ABI (include/tsasr_b200.h); it is loaded with ctypes by _lib.py.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("TSASR_B200_LIB") or os.path.join(_HERE, "libtsasr_b200.so")  # env override: A/B runs of two builds
SOURCES = ["capi.cu", "lattice.cu", "decode.cu", "prep.cu", "predictor.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "177",
]


def _newest_source_mtime():
    m = os.path.getmtime(os.path.join(ROOT, "include", "tsasr_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".inl", ".h")):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def needs_build():
    return not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < _newest_source_mtime()


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> tsasr_b200/libtsasr_b200.so"""
    if not force and not needs_build():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return SO_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
