#!/usr/bin/env python
"""profiles/traffic.json from the committed `ncu --set full` text summaries (profiles/<tag>_ncu_*.txt): per launch
dram__bytes_read.sum + dram__bytes_write.sum and duration of the dominant kernels (bench.py's roofline.traffic).
usage: python tools/traffic_from_summaries.py r2      (k_gicp_linearize is carried over from the capture named in its entry)"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def launches(path, name_re):
    out, cur = [], None
    for ln in open(path):
        m = re.match(r"---- (.*) id (\d+)", ln)
        if m:
            cur = {"kernel": m.group(1), "id": int(m.group(2))} if re.search(name_re, m.group(1)) else None
            if cur:
                out.append(cur)
            continue
        if cur is None:
            continue
        f = ln.split()
        if len(f) >= 3 and f[0] in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            cur[f[0]] = float(f[1]) * UNIT[f[2]]
    return [{"kernel": l["kernel"], "us": round(l["gpu__time_duration.sum"], 2),
             "dram_bytes": int(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"])} for l in out]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    P = os.path.join(ROOT, "profiles")
    old = json.load(open(os.path.join(P, "traffic.json")))
    t = {"source": f"ncu --set full --clock-control none, profiles/{tag}_ncu_*.txt (cold-cache, serialised launches); "
                   "k_gicp_linearize: see its entry"}
    s = launches(os.path.join(P, f"{tag}_ncu_s2m.txt"), r"k_s2m_iteration")
    t["k_s2m_iteration"] = {"launches": s, "dram_bytes_per_launch_last": s[-1]["dram_bytes"], "us_last": s[-1]["us"],
                            "note": "one persistent launch = one whole solve (3 iterations)"}
    b = launches(os.path.join(P, f"{tag}_ncu_s2m_batched.txt"), r"k_s2m_iteration")
    # one iteration of the batched step = a search launch + a fit launch; the steady iteration is the last pair
    t["k_s2m_batched"] = {"launches": b, "dram_bytes_per_launch_last": b[-2]["dram_bytes"] + b[-1]["dram_bytes"],
                          "us_last": round(b[-2]["us"] + b[-1]["us"], 2),
                          "note": "per ITERATION of 256 scans: search launch + fit launch (last pair of the capture); a step is 3-4 iterations"}
    n = launches(os.path.join(P, f"{tag}_ncu_ndt.txt"), r"k_ndt_derivatives")
    t["k_ndt_derivatives"] = {"launches": n, "dram_bytes_per_launch_last": n[-1]["dram_bytes"], "us_last": n[-1]["us"]}
    gp = os.path.join(P, f"{tag}_ncu_gicp.txt")
    if os.path.exists(gp):
        g = launches(gp, r"k_gicp_linearize")
        t["k_gicp_linearize"] = {"launches": g, "dram_bytes_per_launch_last": g[-1]["dram_bytes"], "us_last": g[-1]["us"],
                                 "capture": f"profiles/{tag}_ncu_gicp.txt (the first six evaluations of the 50 M-point align)"}
    else:
        t["k_gicp_linearize"] = old["k_gicp_linearize"]
    # multi-kernel paths: DRAM bytes summed over the launches of ONE call (the capture holds one scan / the first filter)
    sp = os.path.join(P, f"{tag}_ncu_scan.txt")
    if os.path.exists(sp):
        sc = launches(sp, r"k_scan_")[:10]
        t["k_scan_all"] = {"launches": sc, "dram_bytes_per_launch_last": sum(l["dram_bytes"] for l in sc), "us_last": round(sum(l["us"] for l in sc), 2),
                           "note": "sum over the ten k_scan_* launches of one sweep (project + extract_features)"}
    vp = os.path.join(P, f"{tag}_ncu_vx.txt")
    if os.path.exists(vp):
        vx = launches(vp, r"k_vx_|k_rs_|k_scan_apply|k_scan_tile")
        # one filter = from a k_vx_bbox_init to the next one
        starts = [i for i, l in enumerate(vx) if "k_vx_bbox_init" in l["kernel"]]
        one = vx[starts[0]:starts[1]] if len(starts) > 1 else vx
        t["k_vx_all"] = {"launches": one, "dram_bytes_per_launch_last": sum(l["dram_bytes"] for l in one), "us_last": round(sum(l["us"] for l in one), 2),
                         "note": "sum over the launches of one VoxelGrid filter in the capture (1 M points)"}
    cp = os.path.join(P, f"{tag}_ncu_c4.txt")
    if os.path.exists(cp):
        c4 = launches(cp, r"k_gicp_linearize")
        t["k_gicp_linearize_c4"] = {"launches": c4, "dram_bytes_per_launch_last": c4[-1]["dram_bytes"], "us_last": c4[-1]["us"],
                                    "note": "one evaluation of a C4 pair (about 53 k source points)"}
    json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print(json.dumps({k: (v if isinstance(v, str) else {"last": v["dram_bytes_per_launch_last"], "us": v["us_last"]}) for k, v in t.items()}, indent=1))


if __name__ == "__main__":
    main()
