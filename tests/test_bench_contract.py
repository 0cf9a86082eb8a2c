"""Host-side checks of the measurement contract that need no GPU: every `roofline.traffic` figure bench.py reports comes from
profiles/traffic.json (ncu captures summarised by tools/traffic_from_summaries.py), and the summaries it was made from are committed."""
import json
import os
import re

from conftest import ROOT


def test_every_traffic_key_of_the_bench_is_measured():
    src = open(os.path.join(ROOT, "bench.py")).read()
    keys = set(re.findall(r'measured_traffic\("([a-z0-9_]+)"\)', src))
    assert {"k_s2m_iteration", "k_s2m_batched", "k_gicp_linearize", "k_ndt_derivatives"} <= keys
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k in keys:
        assert k in t, k
        e = t[k]
        assert int(e["dram_bytes_per_launch_last"]) > 0 and float(e["us_last"]) > 0 and len(e["launches"]) >= 1
    # the captures the figures come from travel with the repository
    tag = re.search(r"profiles/(\w+)_ncu_", t["source"]).group(1)
    for name in ("s2m", "s2m_batched", "ndt", "gicp", "scan", "vx", "c4"):
        assert os.path.exists(os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.txt")), name


def test_both_arms_share_one_step_definition():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("STEP = ") == 1                       # one constant...
    assert len(re.findall(r'"step": STEP', src)) >= 2       # ...reported by the GPU arm and by the reference arm
