"""Build ``libdiffma_b200.so`` (the C-ABI of include/diffma_b200.h) in-tree with nvcc for sm_100a.

    python -m diffma_b200.build [--force] [--verbose]

The library is plain CUDA C++ (no torch headers): nvcc cross-compiles it without a GPU.  The built ``.so``
lives next to the package (``diffma-diffusion-mamba_b200/lib/``) so it travels with the tree to the GPU box;
it is git-ignored.  ``ensure_built()`` rebuilds only when a source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libdiffma_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "-I", os.path.join(ROOT, "include")]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    return sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(ROOT, "include", "diffma_b200.h")]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _deps())


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libdiffma_b200.so")
    return exe


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile + link under an exclusive file lock: N torchrun ranks of a fresh checkout must not run nvcc into the
    same ``lib/obj/*.o`` concurrently (the first rank builds, the others wait and find a fresh library)."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    cc = nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        extra = os.environ.get("DM_NVCC_EXTRA", "").split()
        cmd = [cc, *ARCH, *FLAGS, *extra, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        results = list(ex.map(compile_one, sources()))
    log = []
    for src, obj, r in results:
        log.append(f"== {os.path.basename(src)}\n{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [cc, *ARCH, "-shared", "-o", LIB + ".tmp", *[o for _, o, _ in results]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(LIB + ".tmp", LIB)
    return LIB


def ensure_built() -> str:
    """Return the library path, building it if it is missing or older than a source.  A failed rebuild RAISES even when
    an old library exists: running a stale ``.so`` after a source edit silently is worse than stopping.  Only a box
    without nvcc (a deployment that ships the prebuilt library) may use the existing file, with a warning."""
    if is_stale():
        have_nvcc = bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")
        if not have_nvcc and os.path.exists(LIB):
            import warnings
            warnings.warn(f"diffma_b200: {LIB} is older than its sources and nvcc is not available; using it as is",
                          RuntimeWarning, stacklevel=2)
            return LIB
        build()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
