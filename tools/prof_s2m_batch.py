"""Batched scan-to-map only (developer tool for ncu): 256 (scan, guess) problems over 4 distinct scans, a few solves."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multi_sensor_slam_tookit_b200 import synth
from multi_sensor_slam_tookit_b200.frontend import ScanFrontEnd
from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer, VoxelGrid
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_input.npz"))
d = {k: z[k] for k in z.files}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ND = 4
scene = synth.CityBlock(synth.MASTER_SEED)
fe, vgc, vgs = ScanFrontEnd(), VoxelGrid(), VoxelGrid()
vgc.setLeafSize(0.2, 0.2, 0.2); vgs.setLeafSize(0.4, 0.4, 0.4)
scans, truths = [], []
for k in range(ND):
    pk = d["pose_truth"].astype(np.float64).copy(); pk[3] += (k - ND // 2) * 2.0 + 0.3
    raw = synth.ring_scan(scene, pk, seed=synth.MASTER_SEED + 7000 + k)
    fe.projectPointCloud(raw, imu=None, deskew=False); f = fe.extractFeatures()
    vgc.setInputCloud(f["corner"]); vgs.setInputCloud(f["surf"])
    scans.append((np.ascontiguousarray(vgc.filter()), np.ascontiguousarray(vgs.filter()))); truths.append(pk.astype(np.float32))
rng = np.random.default_rng(7)
idx = np.arange(B) % ND
poses = np.stack([truths[i] for i in idx]).astype(np.float32)
poses[:, 3:] += rng.uniform(-0.15, 0.15, (B, 3)).astype(np.float32)
poses[:, :3] += np.deg2rad(rng.uniform(-1.0, 1.0, (B, 3))).astype(np.float32)
g = ScanToMapOptimizer(max_batch=B)
g.setInputMap(d["map_corner"], d["map_surf"])
g.setInputScanBatch([scans[i][0] for i in idx], [scans[i][1] for i in idx])
for _ in range(3):
    r = g.scan2MapOptimizationBatch(poses, 4)
    print(g.lastGpuMs(), int(r["iters"].sum()), int(r["converged"].sum()))
