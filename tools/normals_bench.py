"""Developer tool: estimate_normals (BVH build + 30-NN + covariance + eigenvector) on one synthetic C5-style cloud, several
repeats in one process (the first pays the allocations), min / median of the device time.
usage: python tools/normals_bench.py [points] [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools.gicp_bench import c5_clouds_torch
from multi_sensor_slam_tookit_b200 import gicp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
src, tgt, _ = c5_clouds_torch(n, "cuda")
pc = gicp.PointCloud(src)
ms = []
for _ in range(reps):
    pc.estimate_normals(); ms.append(pc.lastGpuMs())
small = gicp.PointCloud(np.ascontiguousarray(np.asarray(src)[:: max(1, n // 53000)][:53000]))
sm = []
for _ in range(reps + 5):
    small.estimate_normals(); sm.append(small.lastGpuMs())
print({"points": n, "ms": [round(m, 2) for m in ms], "min": round(min(ms), 2), "median": round(float(np.median(ms)), 2),
       "small_points": 53000, "small_min_ms": round(min(sm), 3), "small_median_ms": round(float(np.median(sm)), 3)})
