"""Builds libmirres_b200.so (the product: sm_100a CUDA behind the C ABI of include/mirres_b200.h) in-tree.

    python -m mirres_restir_nerf_mesh_b200.build            # product library
    python -m mirres_restir_nerf_mesh_b200.build --hostcheck # test-only host flavour of the per-pixel kernels

nvcc cross-compiles without a GPU.  -fmad=false is part of the numerical contract (include/mirres_fpmath.h).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["bvh_build.cu", "trace.cu", "wave.cu", "env.cu", "restir.cu", "shade.cu", "denoise.cu", "gbuffer.cu", "screen.cu"]
HOSTCHECK_SOURCES = ["trace.cu", "wave.cu", "env.cu", "restir.cu", "shade.cu", "denoise.cu", "gbuffer.cu", "screen.cu"]
LIB = os.path.join(HERE, "libmirres_b200.so")
HOSTCHECK_LIB = os.path.join(HERE, "..", "tests", "_build", "libmirres_hostcheck.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-ccbin", "/usr/bin/g++"] + os.environ.get("MIRRES_NVCC_DEFINES", "").split()  # -D... for tuning experiments only


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    inc = os.path.join(HERE, "..", "include")
    d += [os.path.join(inc, f) for f in os.listdir(inc)]
    d.append(os.path.abspath(__file__))
    return d


def _sources(names):
    return [os.path.join(CSRC, f) for f in names if os.path.exists(os.path.join(CSRC, f))]


def build(force=False, verbose=False):
    if not force and not _stale(LIB, _deps()):
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    for src in _sources(SOURCES):
        obj = os.path.join(HERE, "_obj", os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [NVCC] + ARCH + COMMON + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose:
            print(out)
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd)
    return LIB


def build_hostcheck(force=False):
    """Test-only: the same per-pixel kernel bodies compiled for the host (launches become loops)."""
    target = os.path.abspath(HOSTCHECK_LIB)
    if not force and not _stale(target, _deps()):
        return target
    os.makedirs(os.path.dirname(target), exist_ok=True)
    cmd = [NVCC] + ARCH + ["-O2", "-std=c++17", "-fmad=false", "-DMR_HOST_CHECK", "--expt-relaxed-constexpr", "-ccbin",
                           "/usr/bin/g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp,-msse4.1", "-shared", "-o",
                           target] + _sources(HOSTCHECK_SOURCES) + ["-lgomp"]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    if "--hostcheck" in sys.argv:
        print(build_hostcheck(force=True))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
