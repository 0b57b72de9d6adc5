"""Builds the in-tree CUDA library (nvcc, sm_100a only).  `python -m openvis_b200.build`"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libopenvis_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["ptx.cuh", "gemm_tn.cuh", "prep.cuh", "prep_tma.cuh", "xattn.cuh", "xattn_tc.cuh", "xattn_tc2.cuh", "xattn_tc3.cuh", "chain.cuh", "crop.cuh", "san_attn.cuh", "san_attn_tc.cuh", "postproc.cuh", "msda.cuh", "temporal.cuh", "pixdec.cuh", os.path.join("..", "..", "include", "openvis_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-format-truncation"]


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("NVCC_EXTRA", "").split()          # e.g. -DOVIS_XATTN_TRACE_BUILD (tools/trace_xattn.py)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
