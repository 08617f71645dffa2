"""Build libmono_b200.so (the C ABI of include/mono_abi.h) in-tree with nvcc for sm_100a.

    python fenicsx-beat_b200/build.py [--force]

One translation unit per .cu, compiled in parallel; objects are cached under csrc/build/ and reused
when neither the source nor any header changed.  The .so lands in fenicsx-beat_b200/beat_b200/ so it
travels with the repo snapshot to the GPU box (built artefacts are git-ignored, not gpurun-ignored).
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "beat_b200", "libmono_b200.so")
SOURCES = ["mono_ctx.cu", "ode_kernels.cu", "pde_kernels.cu", "halo.cu", "fem_assemble.cu", "csr_patterns.cu"]  # the last two are host-only C++
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + os.environ.get("MONO_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _header_digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(CSRC, "generated"), os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".h", ".cuh")):
                with open(os.path.join(root, name), "rb") as fh:
                    h.update(name.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    hdr = _header_digest()
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        with open(path, "rb") as fh:
            digest = hashlib.sha256(fh.read() + hdr.encode()).hexdigest()[:16]
        obj = os.path.join(CSRC, "build", src.replace(".cu", f".{digest}.o"))
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append((path, obj))

    def compile_one(job):
        path, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or not os.path.exists(OUT):
        cmd = [nvcc, "-shared", "-o", OUT, *objs, "-ldl", "-lpthread"]  # NCCL is dlopen'ed at mono_comm_init (halo.cu)
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    # drop stale objects
    keep = set(objs)
    for name in os.listdir(os.path.join(CSRC, "build")):
        p = os.path.join(CSRC, "build", name)
        if p not in keep:
            os.remove(p)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
