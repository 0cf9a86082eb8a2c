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


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "b2reg.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def build_lib(force=False, verbose=False, variant=None, variant_flags=()):
    """One object per .cu (compiled in parallel, rebuilt only when the source or a header is newer), then one link.
    No relocatable device code: the translation units only share host functions.
    variant / variant_flags: kernel experiments — a second library variants/libb2reg_<variant>.so built with extra -D flags
    (selected at run time with B2_LIB=...)."""
    if variant is None and not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    base = [nvcc]
    # the image exports CXX=/opt/gcc/bin/g++ whose driver lacks libgomp specs; nvcc only needs a plain host g++
    if os.path.exists("/usr/bin/g++"):
        base += ["-ccbin", "/usr/bin/g++"]
    extra = os.environ.get("B2_NVCC_EXTRA", "").split() + list(variant_flags)
    objdir = os.path.join(HERE, "build" if variant is None else os.path.join("build", variant))
    lib_out = LIB
    if variant is not None:
        os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
        lib_out = os.path.join(HERE, "variants", f"libb2reg_{variant}.so")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    tag = os.path.join(objdir, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(tag) or open(tag).read() != flags_now:
        force = True

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = base + [f for f in NVCC_FLAGS if f != "-shared"] + extra + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    open(tag, "w").write(flags_now)
    cmd = base + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_out] + objs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return lib_out


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
