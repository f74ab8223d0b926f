"""Builds liblds_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m latent_diffusion_speech_b200.build [--force]

The shared object is written next to this file so that it travels to the GPU box with the repo
snapshot; it is git-ignored (source-only history).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "liblds_b200.so")
SOURCES = ["lds_api.cu", "gemm_f32.cu", "gemm_tc.cu", "attention_f32.cu", "attention_tc.cu", "norm.cu", "solver.cu", "vocoder.cu", "units.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build liblds_b200.so")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lds_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, debug_knobs: bool = False) -> str:
    """debug_knobs=True builds liblds_b200_dbg.so with -DLDS_DEBUG_KNOBS (timing experiments that skip stores / epilogues;
    wrong results by construction — never the product library; load it with LDS_B200_LIB=...)."""
    if debug_knobs:
        extra = ["-DLDS_DEBUG_KNOBS"] + os.environ.get("LDS_EXTRA_NVCC_FLAGS", "").split()      # experiment builds only
        return _build(os.path.join(HERE, "liblds_b200_dbg.so"), os.path.join(HERE, "build", "dbg"), extra, verbose)
    if not force and not _stale():
        return LIB_PATH
    return _build(LIB_PATH, os.path.join(HERE, "build"), [], verbose)


def _build(lib_path: str, objdir: str, extra: list, verbose: bool) -> str:
    nvcc = _nvcc()
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode()))
    link = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, debug_knobs="--debug-knobs" in sys.argv))
