"""In-tree build of libb200plan.so (hand-written sm_100a CUDA + C ABI) with nvcc.  No JIT cache: the .so sits next
to this file so it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
TRACE = os.environ.get("B2P_TRACE_BUILD") == "1"   # developer build with in-kernel stage clocks (scripts/tc_trace.py)
PREBUILT = os.environ.get("B2P_LIB_PATH")            # developer A/B: load this prebuilt library as is (never rebuilt)
LIB_PATH = PREBUILT or os.path.join(PKG_DIR, "libb200plan_trace.so" if TRACE else "libb200plan.so")
SOURCES = ["api.cu", "sched.cu", "embed.cu", "conv_ffma.cu", "conv_gemv.cu", "conv_tc.cu", "chain64.cu", "trajpred.cu", "preprocess.cu", "control.cu", "encoder_stem.cu", "encoder_conv.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "--expt-relaxed-constexpr",
] + (["-DB2P_TC_TRACE"] if TRACE else [])


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libb200plan.so cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "b200plan.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if PREBUILT or (not force and not _stale()):
        return LIB_PATH
    objs = []
    nvcc = _nvcc()
    obj_dir = os.path.join(PKG_DIR, "build_trace" if TRACE else "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
