"""Build libragb200.so (sm_100a only) with nvcc, in-tree.

The library lands next to this file as ``libragb200.so`` so that it travels with the
repository snapshot to the GPU box.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
INCLUDE = REPO_ROOT / "include"
LIB_PATH = PKG_DIR / "libragb200.so"
OBJ_DIR = PKG_DIR / "build"

SOURCES = ["common.cu", "select.cu", "bm25.cu", "dense_gemv.cu", "dense_mma.cu", "router.cu", "candidates.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libragb200 cannot be built (there is no fallback)")
    return exe


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "ragb200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path.

    ``profile``: the instrumented diagnostics build (-DRAGB_BM25_PROFILE: per-query / per-phase cycle counters in
    bm25_kernel) as ``libragb200_prof.so`` next to the product library; load it with RAGB_LIB_NAME=libragb200_prof.so
    (scripts/profile_bm25_queries.py does).  Never the default: the counters cost registers."""
    if profile:
        return _build_profile(verbose)
    stamp = OBJ_DIR / "fingerprint"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB_PATH
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = OBJ_DIR / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ_DIR / (src[:-3] + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [str(o) for o, _ in results]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(fp)
    return LIB_PATH


def _build_profile(verbose: bool = False) -> Path:
    out_dir = PKG_DIR / "build_prof"
    out_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DRAGB_BM25_PROFILE"]
    objs = []
    for src in SOURCES:
        obj = out_dir / (src[:-3] + ".o")
        r = subprocess.run([nvcc, *flags, "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(CSRC / src), "-o", str(obj)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        objs.append(str(obj))
    lib = PKG_DIR / "libragb200_prof.so"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *objs],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, profile="--profile" in sys.argv))
