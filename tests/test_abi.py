"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b2reg.h declares.
No compute calls here (there is no GPU in the build container)."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2reg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multi_sensor_slam_tookit_b200 import build, capi
    build.build_lib()
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in b2reg.h but not exported: {missing}"
    assert sorted(capi.SYMBOLS) == syms, "capi.SYMBOLS out of sync with include/b2reg.h"


def test_version_and_error_string_without_gpu():
    from multi_sensor_slam_tookit_b200 import capi
    L = capi.lib()
    assert L.b2_version() >= 100
    assert isinstance(L.b2_last_error(), bytes)
    assert L.b2_device_count() >= 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multi_sensor_slam_tookit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f


def test_header_is_plain_c_and_links():
    """include/b2reg.h is a C header (the reference binds it from C++, Python ctypes and, for the Go / Rust tools of the
    wider toolkit, cgo / FFI): it must compile as C99 and a C program must link against the library."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from multi_sensor_slam_tookit_b200 import capi
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        with open(src, "w") as f:
            f.write('#include "b2reg.h"\n#include <stdio.h>\nint main(void) { printf("%d %d\\n", b2_version(), b2_device_count() >= 0); return 0; }\n')
        exe = os.path.join(d, "t")
        lib_dir = os.path.dirname(capi.LIB_PATH)
        r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), src, "-o", exe,
                            "-L", lib_dir, "-l:" + os.path.basename(capi.LIB_PATH), "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.split()[0] == "100"
