"""In-tree build of libfmm_b200.so (the C-ABI CUDA library) with nvcc for sm_100a.

Usage: ``python -m fall_multimodal_b200.build [--force]``.  The shared object lands in
``fall_multimodal_b200/lib/`` (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libfmm_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-rdc=true",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, hdr_digest: str) -> Path:
    obj = OBJDIR / (src.stem + ".o")
    stamp = OBJDIR / (src.stem + ".stamp")
    key = hashlib.sha256(src.read_bytes() + hdr_digest.encode()).hexdigest()
    if obj.exists() and stamp.exists() and stamp.read_text() == key:
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (OBJDIR / (src.stem + ".ptxas.log")).write_text(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stderr}\n{res.stdout}")
    stamp.write_text(key)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / "build.stamp"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    nvcc = _nvcc()
    h = hashlib.sha256()
    for f in sorted(CSRC.glob("*.cuh")):
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    hdr_digest = h.hexdigest()
    if force:
        for f in OBJDIR.glob("*.stamp"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, hdr_digest), _sources()))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB),
           *map(str, objs), "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stderr}")
    stamp.write_text(dig)
    if verbose:
        for f in sorted(OBJDIR.glob("*.ptxas.log")):
            print(f"== {f.name}")
            print(f.read_text())
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
