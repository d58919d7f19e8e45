"""Build libmuse_b200.so in-tree with nvcc for sm_100a (no torch dependency, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmuse_b200.so")
SOURCES = ["muse_api.cu", "muse_iso_solver.cu", "muse_iso_stream.cu", "muse_draws.cu", "muse_comm.cu", "muse_driver.cu", "muse_outer.cu", "muse_dgemm.cu", "muse_corr.cu", "muse_implicit.cu"]
HEADERS = ["muse_common.cuh", "muse_handle.cuh", "muse_outer_dev.cuh", "muse_group.cuh", "muse_iso_ctl.cuh", "muse_draw_tables.cuh", "muse_normal_math.cuh", os.path.join("..", "..", "include", "muse_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--threads", "0",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmuse_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources into ``museinference.jl_b200/libmuse_b200.so``."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH
