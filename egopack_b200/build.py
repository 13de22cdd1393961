"""Builds ``libegopack_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m egopack_b200.build [--force] [--verbose]

Objects are compiled in parallel into ``egopack_b200/csrc/_build`` and linked into
``egopack_b200/libegopack_b200.so`` (git-ignored; travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libegopack_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("EGP_PTXAS_V") else "-O3"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, refresh=()) -> str:
    """Compile + link.  With an unchanged source digest nothing is rebuilt, except the translation units named in
    ``refresh`` (``__graft_entry__.build()`` always recompiles a few, so "does it build" is answered by the compiler
    and not by a stamp file that travelled with the tree)."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    dig = _digest()
    fresh = os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig
    if not force and fresh and not refresh:
        return LIB
    todo = sources() if (force or not fresh) else [s for s in sources() if s in set(refresh)]
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources()]
    missing = [o for o in objs if not os.path.exists(o)]
    if missing:                                              # objects do not travel in git: build whatever is absent
        with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            list(ex.map(lambda o: _compile(os.path.basename(o)[:-2] + ".cu", verbose), missing))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
