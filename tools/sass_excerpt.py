#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel the resource usage ptxas reports and the mix of memory instructions in the SASS
(128-bit global loads, local-memory spills, shuffles, barriers), plus a short excerpt around the first 128-bit load.
usage: python tools/sass_excerpt.py > profiles/r2_sass_excerpts.txt     (reads the objects of the in-tree build, no GPU needed)"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(HERE, "multi_sensor_slam_tookit_b200", "build")
WANT = [("b2_s2m.o", "k_s2m_iterationILi16ELi2ELi1ELi3"), ("b2_s2m.o", "k_s2m_iterationILi1ELi1ELi1ELi1"), ("b2_s2m.o", "k_s2m_iterationILi1ELi1ELi1ELi2"),
        ("b2_gicp.o", "k_gicp_linearize"), ("b2_grid.o", "k_grid_build_dev"), ("b2_sort.o", "k_rs_onesweep"), ("b2_ndt.o", "k_ndt_derivatives"),
        ("b2_cloud.o", "k_normals"), ("b2_scan.o", "k_scan_features"), ("b2_scan.o", "k_scan_cells")]
PAT = {"LDG.E.128": r"\bLDG\.E\.128", "LDG.E.64": r"\bLDG\.E\.64", "LDG (all)": r"\bLDG\.", "STG.E.128": r"\bSTG\.E\.128", "STG (all)": r"\bSTG\.",
       "LDL (spill loads)": r"\bLDL", "STL (spill stores)": r"\bSTL", "LDS": r"\bLDS", "STS": r"\bSTS", "SHFL": r"\bSHFL", "BAR": r"\bBAR", "ATOM/RED": r"\b(ATOM|RED|ATOMG|ATOMS)\b",
       "DFMA/DADD/DMUL": r"\bD(FMA|ADD|MUL)", "FFMA": r"\bFFMA", "MUFU": r"\bMUFU"}


def main():
    for obj, name in WANT:
        path = os.path.join(OBJ, obj)
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout.splitlines()
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        usage = ""
        for i, ln in enumerate(res):
            if name in ln and i + 1 < len(res):
                usage = res[i + 1].strip()
                break
        m = re.search(r"Function : (\S*%s\S*)\n(.*?)(?=\n\s*Function :|\Z)" % re.escape(name), sass, re.S)
        if not m:
            print(f"==== {name}: not found in {obj}\n")
            continue
        body = [ln for ln in m.group(2).splitlines() if re.search(r"/\*[0-9a-f]{4,6}\*/", ln)]
        print(f"==== {m.group(1)}  ({obj})")
        print(f"  ptxas: {usage}")
        print(f"  SASS instructions: {len(body)} ({len(body) * 16} bytes)")
        for k, p in PAT.items():
            c = sum(1 for ln in body if re.search(p, ln))
            if c:
                print(f"  {k:22s} {c}")
        first = next((i for i, ln in enumerate(body) if re.search(r"\bLDG\.E\.128", ln)), None)
        if first is not None:
            print("  -- excerpt around the first 128-bit global load:")
            for ln in body[max(0, first - 6):first + 10]:
                print("    " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip())
        print()


if __name__ == "__main__":
    main()
