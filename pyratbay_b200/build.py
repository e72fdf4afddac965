"""Build the CUDA C-ABI library in-tree (pyratbay_b200/libpb200_lbl.so) for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box as well as on a B200.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libpb200_lbl.so"
SOURCES = ["engine.cu", "lbl_kernels.cu", "voigt.cu", "microbench.cu",
           "optical_depth.cu", "preprocess.cu", "dense_kernels.cu", "table_ops.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # no implicit FMA contraction in device code: every fused multiply-add is an explicit fma()
    # (the hot gathers), everything else rounds like the reference's C expressions
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "--shared",
]


def lib_path():
    return os.path.join(_HERE, LIB_NAME)


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build the pb200 CUDA library")
    return nvcc


def needs_build():
    out = lib_path()
    if not os.path.exists(out):
        return True
    csrc = os.path.join(_HERE, "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)]
    deps.append(os.path.join(_HERE, "..", "include", "pb200_lbl.h"))
    newest = max(os.path.getmtime(f) for f in deps)
    return os.path.getmtime(out) < newest


def build_library(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into libpb200_lbl.so.  Returns the library path.
    `defines` / `out` build a tuning variant (e.g. defines=["PB200_UNROLL=4"])."""
    variant = out is not None
    out = out or lib_path()
    if not variant and not force and not needs_build():
        return out
    csrc = os.path.join(_HERE, "csrc")
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(csrc, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
