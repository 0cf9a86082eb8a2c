"""Oracle behaviour on the committed workload and the edge cases of the reference path (no GPU)."""
import numpy as np


def test_voxel_grid_properties(oracle):
    rng = np.random.default_rng(11)
    pts = np.zeros((50000, 4), np.float32)
    pts[:, :3] = rng.uniform(-30, 30, (50000, 3)) * np.array([1, 1, 0.1])
    pts[:, 3] = rng.uniform(0, 100, 50000)
    r = oracle.voxel_grid(pts, 0.4)
    out, vop, ovi = r["out"], r["voxel_of_point"], r["out_voxel_idx"]
    assert not r["refused"]
    assert len(out) == len(np.unique(vop))
    assert np.all(np.diff(ovi) > 0)                          # ascending voxel index
    # centroid of each voxel = mean of its members (fp32 tolerance), and it stays inside the voxel's cell
    order = np.argsort(vop, kind="stable")
    sums = np.add.reduceat(pts[order].astype(np.float64), np.flatnonzero(np.diff(np.concatenate([[-1], vop[order]]))), axis=0)
    cnt = np.bincount(np.searchsorted(ovi, vop))
    assert np.allclose(out, sums / cnt[:, None], rtol=1e-5, atol=1e-4)
    assert np.array_equal(np.floor(out[:, :3] * np.float32(1 / 0.4)), np.floor(pts[order][np.cumsum(cnt) - 1, :3] * np.float32(1 / 0.4)))
    # idempotent on its own output when every voxel holds one point
    r2 = oracle.voxel_grid(out, 0.4)
    assert len(r2["out"]) == len(out)


def test_voxel_grid_edge_cases(oracle):
    assert len(oracle.voxel_grid(np.zeros((0, 4), np.float32), 0.2)["out"]) == 0
    one = np.array([[1.5, -2.5, 0.25, 7.0]], np.float32)
    assert np.array_equal(oracle.voxel_grid(one, 0.2)["out"], one)
    # leaf too small for the extent: PCL warns and returns the input unchanged
    far = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2]], np.float32)
    r = oracle.voxel_grid(far, 0.001)
    assert r["refused"] and np.array_equal(r["out"], far)
    # min points per voxel (heading_ws/src/src/PointCloudProcessing.cpp:23-30)
    pts = np.array([[0.01, 0.01, 0.01, 1], [0.02, 0.02, 0.02, 3], [5, 5, 5, 1]], np.float32)
    r = oracle.voxel_grid(pts, 0.1, min_points=2)
    assert len(r["out"]) == 1 and np.allclose(r["out"][0], [0.015, 0.015, 0.015, 2])


def test_c1_scan_to_map_converges_to_truth(oracle, c1):
    s = oracle.Scan2Map(threads=4)
    s.set_map(c1["map_corner"], c1["map_surf"])
    s.set_scan(c1["scan_corner"], c1["scan_surf"])
    r = s.solve(c1["pose_guess"])
    assert r["rc"] == 0 and r["converged"] and 2 <= r["iters"] <= 10
    assert np.all(np.abs(r["pose"][3:] - c1["pose_truth"][3:]) < 0.02)          # metres
    assert np.all(np.abs(r["pose"][:3] - c1["pose_truth"][:3]) < 2e-3)          # radians
    assert not s.get_state()[0]
    assert len(c1["map_corner"]) + len(c1["map_surf"]) == 100000


def test_scan_to_map_guards(oracle, c1):
    s = oracle.Scan2Map(threads=1)
    s.set_map(c1["map_corner"], c1["map_surf"])
    # fewer features than edgeFeatureMinValidNum / surfFeatureMinValidNum: the loop never runs (mapOptmization.cpp:1287)
    s.set_scan(c1["scan_corner"][:10], c1["scan_surf"])
    assert s.solve(c1["pose_guess"])["rc"] == -1
    # fewer than 50 correspondences: LMOptimization returns false, the pose never moves (:1178)
    s.set_scan(c1["scan_corner"][:11], c1["scan_surf"][:101])
    far = c1["pose_guess"].copy(); far[3] += 500
    r = s.solve(far)
    assert r["rc"] == 0 and r["iters"] == 30 and not r["converged"] and np.array_equal(r["pose"], far)


def test_front_end_on_synthetic_scan(oracle):
    from multi_sensor_slam_tookit_b200 import synth
    scene = synth.CityBlock()
    raw = synth.ring_scan(scene, (0, 0, 0.3, 0.0, -24.0, 1.8), seed=5)
    proj = oracle.project(raw, 16, 1800)
    m = len(proj["extracted"])
    assert m == (proj["range_mat"] != np.finfo(np.float32).max).sum() and m > 15000
    # every ring keeps the 5-point margins of cloudExtraction (imageProjection.cpp:580,596)
    per_ring = (proj["range_mat"] != np.finfo(np.float32).max).sum(axis=1)
    cum = np.cumsum(per_ring)
    assert np.array_equal(proj["endRingIndex"], cum - 1 - 5) and np.array_equal(proj["startRingIndex"], cum - per_ring - 1 + 5)
    assert np.all(np.diff(proj["pointColInd"])[np.diff(np.repeat(np.arange(16), per_ring)) == 0] > 0)
    f = oracle.extract_features(proj)
    assert 100 < len(f["corner"]) <= 16 * 6 * 20 and len(f["surf"]) > 1000
    assert np.all(f["label"][f["corner_idx"]] == 1)
    # stable and std::sort modes agree when there are no exact curvature ties among candidates
    f2 = oracle.extract_features(proj, stable=False)
    assert np.array_equal(f["corner_idx"], f2["corner_idx"])
