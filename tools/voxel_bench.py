#!/usr/bin/env python
"""pcl::VoxelGrid at local-map sizes (developer tool for the ncu launch list): 1 M and 4 M points, leaf 0.4 m."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multi_sensor_slam_tookit_b200.registration import VoxelGrid
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_input.npz"))
ms = z["map_surf"]
rng = np.random.default_rng(3)
out = {}
for n in (1_000_000, 4_000_000):
    base = ms[rng.integers(0, len(ms), n)].copy()
    base[:, :3] += rng.normal(0.0, 0.05, (n, 3)).astype(np.float32)
    vg = VoxelGrid(); vg.setLeafSize(0.4, 0.4, 0.4); vg.setInputCloud(base)
    for _ in range(2):
        r = vg.filter()
    dev = 0.0
    for _ in range(3):
        r = vg.filter(); dev += vg.lastGpuMs()
    out[n] = dict(voxels=len(r), gpu_ms=dev / 3)
print(json.dumps(out))
