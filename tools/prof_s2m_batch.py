"""Batched scan-to-map only (developer tool for ncu): 256 pose hypotheses, a few solves."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_input.npz"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(7)
poses = np.tile(d["pose_truth"], (B, 1)).astype(np.float32)
poses[:, 3:] += rng.uniform(-0.15, 0.15, (B, 3)).astype(np.float32)
poses[:, :3] += np.deg2rad(rng.uniform(-1.0, 1.0, (B, 3))).astype(np.float32)
g = ScanToMapOptimizer(max_batch=B)
g.setInputMap(d["map_corner"], d["map_surf"])
g.setInputScanBatch([d["scan_corner"]] * B, [d["scan_surf"]] * B)
for _ in range(3):
    r = g.scan2MapOptimizationBatch(poses, 4)
    print(g.lastGpuMs(), int(r["iters"].sum()))
