"""CPU pins (no GPU) of two restatements the GPU parity tests rely on:
  oracle/o_icp.cpp  o_icp_align  (pcl::IterativeClosestPoint as mapOptmization.cpp:559-586 configures it) against an
                    independent scipy / numpy point-to-point ICP: cKDTree correspondences + the closed-form SVD alignment;
  oracle/pyoracle.fuse_clouds  (PointClouds_Fusion fusion_pointclouds.cpp:55-115) against hand-computed cases.
PCL itself is not installable here; this is what stands between these two restatements and "parity unpinned"."""
import numpy as np

import gicp_cases as G


def _kabsch(src, dst):
    cs, cd = src.mean(0), dst.mean(0)
    H = (src - cs).T @ (dst - cd)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = cd - R @ cs
    return T


def _angle(R):
    """rotation angle from the antisymmetric part (arccos of the trace turns float32 rounding of R into 1e-4 rad)"""
    return float(np.linalg.norm(R - R.T) / (2.0 * np.sqrt(2.0)))


def _numpy_icp(src, tgt, max_d, iters):
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt)
    T = np.eye(4)
    hist = []
    for _ in range(iters):
        cur = src @ T[:3, :3].T + T[:3, 3]
        d, j = tree.query(cur, k=1)
        ok = d <= max_d
        dT = _kabsch(cur[ok], tgt[j[ok]])
        T = dT @ T
        hist.append(dT)                                    # the oracle's history holds the incremental transformations
    return T, hist


def test_icp_align_against_scipy_icp(oracle):
    rng = np.random.default_rng(5)
    tgt = G.lidar_cloud(0, n_rings=16, n_cols=256).astype(np.float32)
    true = G.perturbed(np.eye(4), (0.15, -0.1, 0.05), (0.4, -0.3, 1.0))
    sub = tgt[rng.permutation(len(tgt))[:1500]].astype(np.float64)
    inv = np.linalg.inv(true)
    src = (sub @ inv[:3, :3].T + inv[:3, 3] + rng.normal(0, 0.002, sub.shape)).astype(np.float32)
    # first iterations, step by step: same correspondences (scipy's exact kd-tree) and the same closed-form update
    r = oracle.icp_align(src, tgt, 2.0, 5, 0.0, 0.0)
    assert r["iterations"] == 5
    _, hist = _numpy_icp(src.astype(np.float64), tgt.astype(np.float64), 2.0, 5)
    for k in range(5):
        dT = np.linalg.inv(hist[k]) @ r["history"][k].astype(np.float64)
        assert np.linalg.norm(dT[:3, 3]) < 2e-4 and _angle(dT[:3, :3]) < 2e-5, k
    # to convergence: the generating transform is recovered (noise 2 mm), fitness = mean squared NN distance
    r = oracle.icp_align(src, tgt, 30.0, 100, 1e-6, 1e-6)
    assert r["converged"] and 0 < r["iterations"] < 100
    dT = np.linalg.inv(true) @ r["transformation"].astype(np.float64)
    assert np.linalg.norm(dT[:3, 3]) < 5e-3 and _angle(dT[:3, :3]) < 1e-3
    from scipy.spatial import cKDTree
    T = r["transformation"].astype(np.float64)
    d, _ = cKDTree(tgt.astype(np.float64)).query(src.astype(np.float64) @ T[:3, :3].T + T[:3, 3])
    assert abs(r["fitness_score"] - float((d ** 2).mean())) <= 1e-3 * float((d ** 2).mean()) + 1e-9


def test_fuse_clouds_hand_cases(oracle):
    a = np.array([[1, 2, 3, 10], [0, 0, 0, 11], [-5, 1, 0.5, 12], [np.nan, 0, 0, 13]], np.float32)
    b = np.array([[1, 0, 0, 20], [0, 1, 0, 21]], np.float32)
    T = np.eye(4); T[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]; T[:3, 3] = [10, 20, 30]          # 90 deg about z + shift
    out = oracle.fuse_clouds([a, b], [None, T])
    assert out.shape == (6, 4) and np.array_equal(out[:3], a[:3]) and np.isnan(out[3, 0])
    assert np.array_equal(out[4], [10, 21, 30, 20]) and np.array_equal(out[5], [9, 20, 30, 21])    # intensity carried, order kept
    # external bounds keep min <= v <= max on every axis and drop the non-finite point
    out = oracle.fuse_clouds([a, b], [None, T], ((-1, -1, -1), (10, 21, 30)))
    assert np.array_equal(out[:, 3], [10, 11, 20, 21])                    # limits are inclusive
    out = oracle.fuse_clouds([a, b], [None, T], ((-1, -1, -1), (10, 20.5, 30)))
    assert np.array_equal(out[:, 3], [10, 11, 21])
    # internal bounds keep what lies OUTSIDE the box (any coordinate beyond it), NaN comparisons are false
    out = oracle.fuse_clouds([a], [None], None, ((-1, -1, -1), (2, 3, 4)))
    assert np.array_equal(out[:, 3], [12])
    out = oracle.fuse_clouds([a[:3]], [None], ((-10, -10, -10), (10, 10, 10)), ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)))
    assert np.array_equal(out[:, 3], [10, 12])
    assert oracle.fuse_clouds([], []).shape == (0, 4)
