"""Build A/B variants of libpcs.so (the same engine with different reduction forms, see csrc/gl64.cuh and csrc/ntt.cu)
into build/variants/libpcs_<tag>.so.  The in-tree libpcs.so is not touched.

    python tools/build_variants.py                 # the round-3 A/B set
    python tools/build_variants.py tag=M,S,F,NM,NA # one variant: PCS_REDUCE_FORM, PCS_SQR_FORM, PCS_FOLD_FORM, PCS_NTT_MUL, PCS_NTT_ADD
"""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plonky2_demo_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")
MACROS = ("PCS_REDUCE_FORM", "PCS_SQR_FORM", "PCS_FOLD_FORM", "PCS_NTT_MUL", "PCS_NTT_ADD")
DEFAULT_SET = {
    "base": (0, 0, 0, 0, 0),          # the round-2 build, SASS-identical
    "b_n0": (1, 1, 3, 0, 0), "b_n2": (1, 1, 3, 2, 0),
    "d_n0": (2, 2, 3, 0, 0), "d_n2": (2, 2, 3, 2, 0),
    "h_n0": (1, 2, 3, 0, 0), "h_n2": (1, 2, 3, 2, 0),
}


def build_variant(tag, values):
    d = os.path.join(OUT, tag)
    os.makedirs(d, exist_ok=True)
    defs = [f"-D{m}={v}" for m, v in zip(MACROS, values)]
    objs = []
    for src in B._sources():
        obj = os.path.join(d, src[:-3] + ".o")
        cmd = [B.NVCC] + B.FLAGS + defs + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src} ({tag}):\n{r.stderr}")
        objs.append(obj)
    lib = os.path.join(OUT, f"libpcs_{tag}.so")
    cmd = [B.NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", lib] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"link failed ({tag}):\n{r.stderr}")
    return lib


def main():
    todo = dict(DEFAULT_SET)
    args = [a for a in sys.argv[1:] if "=" in a]
    if args:
        todo = {a.split("=")[0]: tuple(int(x) for x in a.split("=")[1].split(",")) for a in args}
    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        for tag, lib in zip(todo, ex.map(lambda kv: build_variant(*kv), todo.items())):
            print(tag, dict(zip(MACROS, todo[tag])), lib)


if __name__ == "__main__":
    main()
