"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# APJ_B200_LIB selects an experimental build variant by file name (tuning runs only)
LIB = os.path.join(PKG, "lib", os.environ.get("APJ_B200_LIB", "libapj_b200.so"))
SOURCES = ["apj_engine.cu", "apj_step.cu", "apj_rebuild.cu", "apj_observe.cu", "apj_state_io.cu", "apj_setup.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              # no FMA contraction in our own arithmetic: every product/sum rounds like the reference's
              # x86-64 build, so distance predicates and force terms match it bit for bit (CUDA libm
              # routines use explicit fma intrinsics and are unaffected)
              "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "apj_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, defines=(), out=None):
    out = out or LIB
    if not force and out == LIB and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(" ".join(cmd))
    return out


HOST_SRC = os.path.join(PKG, "host", "jam", "jamming.cpp")
HOST_BIN = os.path.join(PKG, "host", "bin", "jam")


def build_host(force=False):
    """The C++ jam driver (host mirror of the reference's Engine / code/classes interface), linked
    against the CUDA library through the C ABI only."""
    deps = [HOST_SRC, os.path.join(PKG, "..", "include", "apj_b200.h")] + [
        os.path.join(PKG, "host", "classes", f) for f in os.listdir(os.path.join(PKG, "host", "classes"))]
    if not force and os.path.exists(HOST_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_BIN) for d in deps):
        return HOST_BIN
    build_library()
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-o", HOST_BIN, HOST_SRC,
           "-L" + os.path.dirname(LIB), "-lapj_b200", "-Wl,-rpath,$ORIGIN/../../lib"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("host build failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
    return HOST_BIN


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_host(force=True))
