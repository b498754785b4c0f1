"""Build libcarel_head.so (sm_100a only) in-tree with nvcc.

    python -m carel_vae_b200.build [--force]

Every .cu under csrc/ is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
and linked into carel_vae_b200/libcarel_head.so.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libcarel_head.so"
OBJ = PKG / "csrc" / "_obj"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-DCAREL_NO_FAST_MATH",
] + os.environ.get("CAREL_EXTRA_FLAGS", "").split()


def _newest_header() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "carel_head.h"]
    return max(h.stat().st_mtime for h in hs)


def _compile(src: Path, force: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, _newest_header()):
        return obj
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    srcs = sorted(CSRC.glob("*.cu"))
    OBJ.mkdir(exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or not OUT.exists() or any(o.stat().st_mtime > OUT.stat().st_mtime for o in objs):
        cmd = [NVCC, "-shared", "-o", str(OUT), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"built {OUT}")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
