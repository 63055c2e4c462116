"""Build libpcs.so (the C-ABI engine) in-tree with nvcc for sm_100a.

    python -m plonky2_demo_b200.build [--force]

Every .cu under csrc/ is compiled to build/*.o in parallel (the Poseidon kernels are fully
unrolled and take ~1 min each) and linked into plonky2_demo_b200/libpcs.so.  The .so is
git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpcs.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -ccbin: the image exports CC/CXX=/opt/gcc wrappers; use the system g++ explicitly
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "pcs.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, extra):
    obj = os.path.join(BUILD, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), _headers_mtime()):
        return obj, False
    cmd = [NVCC] + FLAGS + extra + ["-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build(force=False, verbose=False, extra=()):
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, list(extra)), srcs))
    objs = [o for o, _ in res]
    changed = any(c for _, c in res)
    if changed or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(("built " if changed else "up to date: ") + LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
