"""
Builds libgib200.so (the C-ABI CUDA library of the hot path) in-tree with nvcc for sm_100a.

    python -m graphinvent_b200.build            # incremental
    python -m graphinvent_b200.build --force

The .so lives at graphinvent_b200/lib/libgib200.so: git-ignored, but it travels to the GPU box
with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgib200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_dep():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def is_stale():
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_dep()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and not is_stale():
        return LIB
    # every rank of a torchrun job imports the package: serialise concurrent builds of the same tree
    import fcntl
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    dep_t = _newest_dep()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= dep_t:
        return LIB          # another process built it while this one waited for the lock
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC} and {LIB} is missing or stale")

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep_t:
            return obj
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + ".tmp"
    r = subprocess.run([NVCC, "-shared", "-o", tmp, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt",
                        "-L/usr/local/cuda/lib64"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
