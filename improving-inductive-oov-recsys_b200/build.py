"""Build the C-ABI library (liboov_b200.so) in-tree with nvcc for sm_100a.

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
`python -m ...build` is not needed: call `build()` (done by `__graft_entry__.build()`).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liboov_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(os.path.dirname(PKG_DIR), "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _headers() -> list[str]:
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(os.path.dirname(PKG_DIR), "include", "*.h"))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ (one object per file, stale ones only, in parallel) and link them into one shared
    library.  Raises on failure."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build liboov_b200.so (no CPU fallback exists)")
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = max([os.path.getmtime(h) for h in _headers()] + [os.path.getmtime(__file__)])
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    compile_flags += os.environ.get("OOV_NVCC_EXTRA", "").split()       # e.g. -DOOV_LSH_TRACE for scripts/trace_lsh.py

    def compile_one(src: str):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, 0, ""
        proc = subprocess.run([nvcc] + compile_flags + ["-c", src, "-o", obj], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        return obj, proc.returncode, proc.stdout

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    failed = [out for _, rc, out in results if rc != 0]
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    if verbose:
        print("\n".join(out for _, _, out in results if out))
    link = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH]
                          + [o for o, _, _ in results], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
