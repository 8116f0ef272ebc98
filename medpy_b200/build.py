"""In-tree build of the native pieces (no JIT cache: the built files travel with the repo snapshot).

  medpy_b200/lib/libmedpy_b200_gc.so      C ABI + CUDA kernels   (nvcc, sm_100a only)
  medpy_b200/_mgc<EXT_SUFFIX>             pybind11 binding       (g++, links the C-ABI library via $ORIGIN/lib)

``python -m medpy_b200.build`` or ``medpy_b200.build.build_all()``; __graft_entry__.build() calls the latter.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmedpy_b200_gc.so")
EXT = os.path.join(HERE, "_mgc" + sysconfig.get_config_var("EXT_SUFFIX"))
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # IEEE everywhere: denormals kept, exact div/sqrt, and no FMA contraction so the float64 weights equal
    # numpy's
    "-ftz=false", "-prec-div=true", "-prec-sqrt=true", "-fmad=false",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_lib(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(INCLUDE, "medpy_b200_graphcut.h"))
    if not force and not _newer(LIB, srcs):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", LIB,
                                                                          os.path.join(CSRC, "gc_api.cu")]
    subprocess.check_call(cmd)
    return LIB


def build_ext(force=False):
    src = os.path.join(CSRC, "gc_pybind.cpp")
    if not force and not _newer(EXT, [src, LIB, os.path.join(INCLUDE, "medpy_b200_graphcut.h")]):
        return EXT
    import pybind11
    cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I", sysconfig.get_paths()["include"], "-I", pybind11.get_include(), "-I", INCLUDE,
           src, "-o", EXT, "-L", LIBDIR, "-lmedpy_b200_gc", "-Wl,-rpath,$ORIGIN/lib"]
    subprocess.check_call(cmd)
    return EXT


def build_all(force=False, verbose=False):
    build_lib(force=force, verbose=verbose)
    build_ext(force=force)
    return LIB, EXT


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    print(EXT)
