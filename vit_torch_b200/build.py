"""Build libvitk.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m vit_torch_b200.build [--force] [--verbose]

The shared library has no torch dependency: it exports the plain-C entry points declared in include/vitk.h.
Objects are rebuilt only when a source or header is newer than the object (nvcc cross-compiles without a GPU).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
# VITK_NVCC_EXTRA="-DVITK_TRACE" builds an instrumented copy (libvitk_dbg.so, objects in build_dbg/) next to the
# product library; load it with VITK_LIB=.../libvitk_dbg.so (scripts/trace_attn.py).
EXTRA = os.environ.get("VITK_NVCC_EXTRA", "").split()
LIB = os.path.join(HERE, "libvitk_dbg.so" if EXTRA else "libvitk.so")
OBJ_DIR = os.path.join(HERE, "build_dbg" if EXTRA else "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
] + EXTRA


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "vitk.h"))
    return hs


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = _headers()
    # the library travels to the GPU box without its object files: up to date = newer than every source and header
    if not force and not _stale(LIB, sources() + hdrs):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        log = p.stdout + p.stderr
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
