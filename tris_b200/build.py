"""Build libtris_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tris_b200.build            # incremental
    python -m tris_b200.build --force
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(os.path.dirname(HERE), "build", "obj")
LIB = os.path.join(LIBDIR, "libtris_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v"]
if os.environ.get("TRIS_GEMM_MAXNREG"):      # experiment knob: cap the GEMM kernel's registers (co-residency with BN kernels)
    FLAGS.append("-DTRIS_GEMM_MAXNREG=" + os.environ["TRIS_GEMM_MAXNREG"])


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "tris_sm100.h"))
    stamp = os.path.join(LIBDIR, ".build_digest")
    dg = _digest(srcs + hdrs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dg:
        return LIB
    hd = _digest(hdrs)

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        tag = obj + ".digest"
        d = _digest([src]) + hd
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == d:
            return obj, ""
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        open(tag, "w").write(d)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(compile_one, srcs))
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    objs = [o for o, _ in res]
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dg)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
