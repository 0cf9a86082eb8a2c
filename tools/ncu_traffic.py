#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch of the
dominant kernels (bench.py's roofline.traffic). Read here, no GPU needed.
usage: python tools/ncu_traffic.py <tag>      (reads gpurun_out/<tag>_prof_{s2m,gicp,ndt}.ncu-rep)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, k):
        i = hdr.index(k)
        try:
            return float(r[i]) * UNIT.get(units[i], 1.0)
        except ValueError:
            return float("nan")
    out = []
    for r in rows[2:]:
        t = val(r, "gpu__time_duration.sum")
        tu = units[hdr.index("gpu__time_duration.sum")]
        t_us = float(r[hdr.index("gpu__time_duration.sum")]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1.0)
        out.append({"kernel": r[hdr.index("Kernel Name")].split("(")[0], "us": t_us,
                    "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")})
    return out


def main():
    tag = sys.argv[1]
    res = {"source": f"ncu --set full --clock-control none, captures {tag}_prof_*.ncu-rep (summaries in profiles/{tag}_ncu_*.txt)"}
    for key, name in (("k_s2m_iteration", "s2m"), ("k_gicp_linearize", "gicp"), ("k_ndt_derivatives", "ndt")):
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{name}.ncu-rep")
        if not os.path.exists(rep):
            continue
        ls = [l for l in launches(rep) if l["dram_bytes"] == l["dram_bytes"]]
        if not ls:
            continue
        res[key] = {"launches": [{"us": round(l["us"], 2), "dram_bytes": int(l["dram_bytes"])} for l in ls],
                    "dram_bytes_per_launch_last": int(ls[-1]["dram_bytes"]), "us_last": round(ls[-1]["us"], 2)}
    json.dump(res, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
