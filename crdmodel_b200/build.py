"""Build the two libraries in-tree:
  libcrd_b200.so  sm_100a CUDA kernels + the C ABI of include/crd_b200.h (nvcc)
  libcrd_ark.so   the explicit RK driver behind the ARKode-legacy names and the generic N_VXxx dispatchers
                  (include/crd_sundials_compat.h, crd_ark.h; host-only, g++).  Kept OUT of libcrd_b200.so so that a program
                  which links the real SUNDIALS 2.x (option A) never sees two definitions of ARKode /
                  N_VLinearSum; programs without SUNDIALS (option B, the drivers here) link both.

    python -m crdmodel_b200.build        # or crdmodel_b200.build.build()

nvcc cross-compiles without a GPU.  The library lands in crdmodel_b200/lib/ (git-ignored, travels
to the GPU box with the snapshot).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libcrd_b200.so")
LIB_ARK = os.path.join(LIB_DIR, "libcrd_ark.so")

CUDA_SOURCES = ["csrc/crd_ctx.cu", "csrc/crd_rhs.cu", "csrc/crd_nvector.cu", "csrc/crd_resident.cu", "csrc/crd_snapshot.cu"]
HOST_SOURCES = ["host/crd_ark.cpp", "host/crd_nvector_generic.c"]
HEADERS = ["csrc/crd_common.cuh", "csrc/crd_grid.cuh", "csrc/crd_rhs_kernels.cuh", "csrc/crd_rhs_pair.cuh", "csrc/crd_rhs_point.cuh", "csrc/crd_fused.cuh", "csrc/crd_tables.hpp", "../include/crd_b200.h",
           "host/crd_pow.h", "../include/crd_ark.h", "../include/crd_sundials_compat.h"]
NVCC_FLAGS = ["-O3", "-std=c++17", "--threads", "4", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xlinker", "-Bsymbolic", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in sources)


def needs_build():
    return _stale(LIB, CUDA_SOURCES + HEADERS) or _stale(LIB_ARK, HOST_SOURCES + HEADERS[-3:])


def build(force=False, verbose=False, profiling_variants=False):
    """profiling_variants: also compile the tilings that measured slower (-DCRD_PROFILING_VARIANTS; profiles/README.md)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]
    if force or profiling_variants or _stale(LIB, CUDA_SOURCES + HEADERS):
        cmd = [_nvcc()] + NVCC_FLAGS + inc + (["-DCRD_PROFILING_VARIANTS"] if profiling_variants else []) + \
              ["-o", LIB] + [os.path.join(HERE, s) for s in CUDA_SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    if force or _stale(LIB_ARK, HOST_SOURCES + HEADERS[-3:]):
        objs = []
        for src in HOST_SOURCES:
            obj = os.path.join(LIB_DIR, os.path.basename(src) + ".o")
            cc = ["g++", "-std=c++17"] if src.endswith(".cpp") else ["gcc", "-std=c99"]
            subprocess.run(cc + ["-O2", "-ffp-contract=off", "-fPIC", "-c", os.path.join(HERE, src), "-o", obj] + inc, check=True)
            objs.append(obj)
        cmd = ["g++", "-shared", "-Wl,-Bsymbolic", "-o", LIB_ARK] + objs + ["-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


def build_drivers(verbose=False):
    """The four host executables (same names as the reference's CMake targets, CMakeLists.txt:37-50)."""
    src = os.path.join(HERE, "host", "crd_driver.cpp")
    if not os.path.exists(src):
        return []
    out = []
    os.makedirs(os.path.join(ROOT, "bin"), exist_ok=True)
    for name, model in (("FHNmodel_torus", 0), ("GoldbeterModel_torus", 1), ("FHNmodel_flat", 2), ("GoldbeterModel_flat", 3)):
        exe = os.path.join(ROOT, "bin", name)
        deps = [src, LIB, LIB_ARK] + [os.path.join(HERE, "host", h) for h in ("crd_ini.hpp", "crd_writer.hpp", "crd_workers.hpp", "crd_steady.hpp")]
        if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(d) for d in deps):
            out.append(exe)
            continue
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-DCRD_DRIVER_MODEL=%d" % model,
               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "host"), src, "-o", exe,
               "-L" + LIB_DIR, "-lcrd_ark", "-lcrd_b200", "-Wl,-rpath,$ORIGIN/../crdmodel_b200/lib", "-lpthread", "-lrt"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        out.append(exe)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, profiling_variants="--profiling-variants" in sys.argv)
    build_drivers(verbose=True)
    print(LIB)
