"""Build libcamc2v_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m camc2v_b200.build [--force]

The shared object exposes only the C ABI of include/camc2v_b200.h, has no torch / Python dependency
and is loaded through ctypes by camc2v_b200._lib.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcamc2v_b200.so")
OUT_FP16 = os.path.join(HERE, "libcamc2v_b200_fp16.so")        # same kernels with IEEE-half operands (-DC2V_OPERAND_FP16)
OBJ = os.path.join(HERE, "csrc", "_obj")
SOURCES = ["api.cu", "gemm_tc.cu", "gemm_ps.cu", "attn_fa.cu", "epi_maps.cu", "attn_small.cu", "attn_t16.cu", "norm.cu", "elementwise.cu", "epipolar.cu", "pose.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"] + os.environ.get("C2V_NVCC_EXTRA", "").split()


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, fp16: bool = False) -> str:
    out = OUT_FP16 if fp16 else OUT
    obj_dir = os.path.join(HERE, "csrc", "_obj_fp16" if fp16 else "_obj")
    extra = ["-DC2V_OPERAND_FP16"] if fp16 else []
    return _build(out, obj_dir, extra, force, verbose)


def _build(OUT: str, OBJ: str, extra, force: bool, verbose: bool) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "camc2v_b200.h"))
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        run([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return OUT


if __name__ == "__main__":
    if "--fp16" in sys.argv or "--all" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, fp16=True))
    if "--fp16" not in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
