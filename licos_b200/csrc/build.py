"""Builds liblicos_b200.so (sm_100a) in-tree with plain nvcc.  No torch headers are involved: the
library is a pure C ABI (include/licos_b200.h); Python reaches it through ctypes."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "lib", "liblicos_b200.so")
SOURCES = ["entropy.cu", "conv_engine.cu", "rans_device.cu", "metrics.cu", "train_kernels.cu", "federated_nccl.cu",
           "host_codec.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-pthread",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "licos_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(PKG, "lib", os.path.splitext(src)[0] + ".o")
        cmd = [NVCC, *FLAGS, *os.environ.get("LICOS_NVCC_EXTRA", "").split(), "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if out.strip() and (verbose or pr.returncode != 0):
            print(f"--- {src}\n{out}", file=sys.stderr)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building liblicos_b200.so")
    link = [NVCC, "-shared", "-o", OUT, *objs, "-lcudart", "-ldl", "-Xcompiler", "-pthread"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
