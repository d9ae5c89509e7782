"""Build the C-ABI CUDA library in-tree: flid_b200/libflid_b200.so (sm_100a only).

    python -m flid_b200.build [--force]

nvcc cross-compiles without a GPU; the CUDA runtime is linked statically so the
library loads (and exports its symbols) on a CPU-only box too.
"""
import concurrent.futures as cf
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libflid_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
BASE_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
FLAGS = BASE_FLAGS + ["-I", INCLUDE, "-I", CSRC]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for f in sorted(os.listdir(d)):
            p = os.path.join(d, f)
            if os.path.isfile(p) and f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(BASE_FLAGS).encode())      # no absolute paths: the tree is copied to other boxes
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def _up_to_date(stamp, digest):
    return os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read() == digest


def build(force=False, verbose=True):
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest()
    if not force and _up_to_date(stamp, digest):
        return LIB
    # several ranks of one job may get here at once: build under an exclusive lock, re-check inside
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(stamp, digest):
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp, digest, verbose):
    if not os.path.isfile(NVCC):
        if os.path.isfile(LIB):
            return LIB  # GPU box without a toolkit: use the shipped binary
        raise RuntimeError(f"nvcc not found at {NVCC} and {LIB} is missing")
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    if verbose:
        print(f"[flid_b200.build] nvcc sm_100a: {', '.join(srcs)}", file=sys.stderr)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    tmp_lib = LIB + ".tmp"
    cmd = [NVCC, "-shared", "-o", tmp_lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_lib, LIB)                      # atomic: a concurrent loader never sees a half-written file
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
