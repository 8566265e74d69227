"""In-tree build of libgsttaco.so with nvcc for sm_100a (no JIT cache: the built .so travels with
the source tree)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libgsttaco.so")
# flipped to True once the bf16 tcgen05 decoder is the default throughput path of bench.py
BF16_READY = True

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libgsttaco.so cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, "api.cu")]


def _newest_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for dirpath, _, files in os.walk(root):
            for f in files:
                if f.endswith((".cu", ".cuh", ".h")):
                    m = max(m, os.path.getmtime(os.path.join(dirpath, f)))
    return m


def needs_build() -> bool:
    return (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < _newest_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + sources() + ["-o", LIB_PATH]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libgsttaco.so")
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
