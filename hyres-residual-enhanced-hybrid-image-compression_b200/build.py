"""In-tree build of the C-ABI library (``csrc/libhyres_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the repo snapshot.  Called by ``__graft_entry__.build()``.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libhyres_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-pthread", "--use_fast_math=false",
]


def _sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp"))
    )


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ) + [os.path.join(ROOT, "include", "hyres_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())  # not the path: the snapshot on the GPU box lives elsewhere
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu/.cpp under csrc/ into one shared library."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    for src in _sources():
        obj = os.path.splitext(src)[0] + ".o"
        cmd = [nvcc, *flags, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", src, "-o", obj]
        if src.endswith(".cpp"):
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lpthread"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
