"""In-tree build of libb2reg.so (hand-written CUDA for sm_100a behind the C ABI of include/b2reg.h).

nvcc cross-compiles without a GPU. The library is built next to this file so it travels with the repo snapshot
to the GPU box; there is no JIT cache and no pip install.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2reg.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # bit-parity with the reference's no-FMA x86-64 build
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b2reg.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + sources()
    extra = os.environ.get("B2_NVCC_EXTRA")
    if extra:
        cmd += extra.split()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ whose driver lacks libgomp specs; nvcc only needs a plain host g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
