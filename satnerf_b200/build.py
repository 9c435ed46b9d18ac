"""Builds satnerf_b200/libsatnerf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m satnerf_b200.build [--force]

The library has a plain C ABI (include/satnerf_b200.h) and links the CUDA runtime statically, so it
has no dependency on torch; satnerf_b200/capi.py binds it with ctypes.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsatnerf_b200.so")
DEV_LIB = os.path.join(PKG, "libsatnerf_b200_dev.so")       # product sources + -DSNB_DEV_BUILD + microbenchmarks (include/satnerf_b200_dev.h)
SOURCES = ["layout.cu", "sampling.cu", "composite.cu", "simt_field.cu", "tc_field.cu", "tc_backward.cu", "tc_bwd.cu", "geo.cu", "optim.cu", "capi.cu"]
DEV_SOURCES = SOURCES + ["mma_rate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "satnerf_b200.h")]
    return _newest(deps) > os.path.getmtime(LIB)


def _compile_and_link(nvcc, sources, objdir, lib, extra, verbose):
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in sources:
        if not os.path.exists(os.path.join(CSRC, s)):
            continue
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")


def build_library(force=False, verbose=False, dev=True):
    """Builds the product library (and, with dev=True, the developer library with the debug entry points)."""
    if not force and not needs_build() and (not dev or os.path.exists(DEV_LIB)):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    _compile_and_link(nvcc, SOURCES, os.path.join(PKG, "build"), LIB, [], verbose)
    if dev:
        _compile_and_link(nvcc, DEV_SOURCES, os.path.join(PKG, "build", "dev"), DEV_LIB, ["-DSNB_DEV_BUILD"], False)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
