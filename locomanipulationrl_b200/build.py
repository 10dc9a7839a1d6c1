"""Builds liblm_b200.so (the C-ABI library, include/lm_b200.h) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA runtime code behind `extern "C"`.
Run ``python -m locomanipulationrl_b200.build`` or call :func:`build`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liblm_b200.so")
STAMP = os.path.join(PKG_DIR, "liblm_b200.stamp")
SOURCES = ["lm_task.cu", "lm_pd.cu", "lm_gae.cu", "lm_mlp.cu", "lm_gnn.cu", "lm_gnn_tc.cu", "lm_rollout.cu", "lm_ppo.cu", "lm_noise.cu", "lm_task_cc.cu", "lm_mlp_train.cu", "lm_comm.cu", "lm_gnn_train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC,-O2", "-cudart", "static", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=...)")
    return cand


def _digest(paths: List[str]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def source_files() -> List[str]:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return srcs


def build_variant(out_path: str, defines: List[str]) -> str:
    """Experiment builds (-D switches of the kernels' diagnostics) next to the product library; never stamped."""
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out_path] + source_files()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + out_path)
    return out_path


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = source_files()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "lm_b200.h"))
    digest = _digest(deps)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building liblm_b200.so")
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
