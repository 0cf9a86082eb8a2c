"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the CPU oracle cannot follow here):
  C5  50 M x 50 M map-to-map GICP: sharded sums add up to the unsharded ones (the identity the NCCL all-reduce relies on),
      evaluations are bit-reproducible, the seeded search returns the same transformation bits as the unseeded one,
      the registration recovers the generating transform;
  C1  256-scan batch: every scan of the batch ends where the single-scan solve of the same problem ends.
Inputs are generated from fixed seeds on the device (tools/gicp_bench.c5_clouds_torch, the committed C1 golden input)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c5_50m_properties(b2):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a 60 GB device")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gicp_bench as GB
    from multi_sensor_slam_tookit_b200 import capi, gicp
    n = 50_000_000
    src, tgt, T_true = GB.c5_clouds_torch(n, "cuda")
    torch.cuda.empty_cache()
    sp, tp = gicp.PointCloud(src), gicp.PointCloud(tgt)
    del src, tgt
    sp.estimate_normals(); tp.estimate_normals()
    g = gicp.GeneralizedICP(1.0, 0.005, -1.0, -1.0, 6)           # negative thresholds: six fixed iterations
    g.setInputTarget(tp); g.setInputSource(sp)
    L = capi.lib()
    # (1) block-cyclic shards: per-rank sums add up to the whole (what the 30-double all-reduce computes); counts exactly
    T0 = np.eye(4)
    full = g.linearize(T0)
    assert np.array_equal(full, g.linearize(T0))                 # bit-reproducible
    for world in (2, 8):
        parts = np.zeros(30)
        pts = 0
        for rank in range(world):
            capi.check(L.b2_gicp_set_shard(g._h, rank, world, None))
            parts += g.linearize(T0)
            pts += g.indexInfo()["shard_points"]
        assert pts == n
        assert parts[27] == full[27]
        assert np.abs(parts - full).max() <= 1e-11 * np.abs(full).max()
    capi.check(L.b2_gicp_set_shard(g._h, 0, 1, None))
    assert 0.2 * n < full[27] <= n
    # (2) the previous correspondences only narrow the search: same bits with and without them
    seeded = g.align(np.eye(4), want_correspondences=False)
    os.environ["B2_GICP_NO_SEED"] = "1"
    try:
        plain = g.align(np.eye(4), want_correspondences=False)
    finally:
        del os.environ["B2_GICP_NO_SEED"]
    assert np.array_equal(seeded.transformation, plain.transformation)
    assert seeded.fitness == plain.fitness and seeded.inlier_rmse == plain.inlier_rmse and seeded.iterations == plain.iterations == 6
    # (3) it registers: 0.3 m / 0.8 deg off at the start, noise 0.02 m per coordinate on both clouds
    dT = np.linalg.inv(T_true) @ seeded.transformation
    assert np.linalg.norm(dT[:3, 3]) < 2e-3 and np.arccos(np.clip((np.trace(dT[:3, :3]) - 1) / 2, -1, 1)) < 1e-5
    assert seeded.fitness > 0.999 and 0.03 < seeded.inlier_rmse < 0.06
    at_truth = g.linearize(T_true)
    assert at_truth[27] > 0.999 * n and abs(np.sqrt(at_truth[28] / at_truth[27]) - seeded.inlier_rmse) < 2e-3
    del g, sp, tp
    L.b2_trim_memory()


def test_c1_batch_256_matches_single_scan(b2):
    from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer
    d = np.load(os.path.join(ROOT, "tests", "golden", "c1_input.npz"))
    B = 256
    rng = np.random.default_rng(7)
    poses = np.tile(d["pose_truth"], (B, 1)).astype(np.float32)
    poses[:, 3:] += rng.uniform(-0.15, 0.15, (B, 3)).astype(np.float32)
    poses[:, :3] += np.deg2rad(rng.uniform(-1.0, 1.0, (B, 3))).astype(np.float32)
    gb = ScanToMapOptimizer(max_batch=B)
    gb.setInputMap(d["map_corner"], d["map_surf"])
    gb.setInputScanBatch([d["scan_corner"]] * B, [d["scan_surf"]] * B)
    rb = gb.scan2MapOptimizationBatch(poses.copy(), 30)
    assert rb["converged"].all() and rb["iters"].max() <= 6
    # the throughput shape (thread per feature, bounded search) against the latency shape, problem by problem
    g1 = ScanToMapOptimizer()
    g1.setInputMap(d["map_corner"], d["map_surf"])
    g1.setInputScan(d["scan_corner"], d["scan_surf"])
    for b in range(0, B, 17):
        g1.transformTobeMapped = poses[b].copy()
        r1 = g1.scan2MapOptimization(30, want_matP=False)
        assert r1["iters"] == rb["iters"][b]
        dp = np.abs(np.asarray(g1.transformTobeMapped, np.float64) - rb["poses"][b].astype(np.float64))
        assert dp[:3].max() <= 1e-6 and dp[3:].max() <= 1e-5      # north_star tolerance: 1e-6 rad / 1e-5 m
    # all hypotheses end at the same optimum (the scan was generated at pose_truth)
    spread = rb["poses"].astype(np.float64) - rb["poses"].astype(np.float64).mean(0)
    assert np.abs(spread[:, 3:]).max() < 5e-3 and np.abs(spread[:, :3]).max() < 5e-4
