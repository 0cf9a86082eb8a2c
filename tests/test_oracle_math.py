"""Pins the oracle's restated library routines against what is importable here:
cv2 4.13 for cv::eigen / cv::solve(DECOMP_QR) / cv::invert / gemm, numpy for least squares, brute force for kNN."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def test_jacobi_eigen3_bit_exact_vs_cv2(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(500):
        B = rng.normal(size=(5, 3)).astype(np.float32) * rng.uniform(0.01, 2)
        B[:, rng.integers(0, 3)] *= rng.uniform(0.01, 5)
        A = (B.T @ B / 5).astype(np.float32)
        A = ((A + A.T) / 2).astype(np.float32)
        W, V = np.zeros(3, np.float32), np.zeros(9, np.float32)
        L.o_eigen3(A.reshape(-1).copy(), W, V)
        ok, w, v = cv2.eigen(A)
        assert np.array_equal(w.reshape(-1), W) and np.array_equal(v.reshape(-1), V)


def test_eigen6_qr_solve_lu_invert_bit_exact_vs_cv2(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for _ in range(300):
        J = rng.normal(size=(200, 6)).astype(np.float32) * np.array([30, 30, 30, 1, 1, 1], np.float32)
        A = (J.T.astype(np.float64) @ J.astype(np.float64)).astype(np.float32)
        b = (J.T.astype(np.float64) @ rng.normal(size=200)).astype(np.float32)
        W, V = np.zeros(6, np.float32), np.zeros(36, np.float32)
        L.o_eigen6(A.reshape(-1).copy(), W, V)
        ok, w, v = cv2.eigen(A)
        assert np.array_equal(w.reshape(-1), W) and np.array_equal(v.reshape(-1), V)
        x = np.zeros(6, np.float32)
        assert L.o_qr_solve6(A.reshape(-1).copy(), b, x) == 1
        ok, xc = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_QR)
        assert np.array_equal(xc.reshape(-1), x)
        Vi = np.zeros(36, np.float32)
        L.o_lu_invert6(np.ascontiguousarray(v).reshape(-1), Vi)
        r, vic = cv2.invert(v)
        assert np.array_equal(vic.reshape(-1), Vi)


def test_gemm_double_accumulation_matches_cv2():
    # matAtA = matAt * matA on CV_32F: products and sums in double, one rounding (mapOptmization.cpp:1225)
    rng = np.random.default_rng(3)
    A = (rng.normal(size=(3000, 6)) * np.array([30, 30, 30, 1, 1, 1])).astype(np.float32)
    ref = cv2.gemm(np.ascontiguousarray(A.T), A, 1.0, None, 0.0)
    mine = (A.T.astype(np.float64) @ A.astype(np.float64)).astype(np.float32)
    assert np.max(np.abs(ref - mine) / np.abs(mine)) < 2e-7     # same up to the last float bit


def test_plane_fit_matches_lstsq(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(4)
    for _ in range(300):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        d = rng.uniform(2, 30)
        p = rng.normal(size=(5, 3)) * 0.4
        p -= np.outer(p @ n, n)
        p += n * d + rng.normal(size=(5, 3)) * 0.01
        A = p.astype(np.float32)
        x = np.zeros(3, np.float32)
        L.o_plane5(A.reshape(-1).copy(), x)
        ref = np.linalg.lstsq(A.astype(np.float64), -np.ones(5), rcond=None)[0]
        assert np.allclose(x, ref, rtol=2e-3, atol=1e-5)


def test_kdtree_equals_brute_force(oracle):
    rng = np.random.default_rng(5)
    pts = np.zeros((20000, 4), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, (20000, 3))
    pts[:500, :3] = pts[500:1000, :3]                       # exact duplicates: ties must fall to the smaller index
    q = np.zeros((3000, 4), np.float32)
    q[:, :3] = rng.uniform(-22, 22, (3000, 3))
    q[:100, :3] = pts[:100, :3]
    i1, d1 = oracle.knn(pts, q, 5)
    i2, d2 = oracle.knn(pts, q, 5, brute=True)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert np.all(np.diff(d1, axis=1) >= 0)


def test_kdtree_matches_scipy(oracle):
    scipy_spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(6)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.normal(size=(5000, 3)) * 10
    q = np.zeros((500, 4), np.float32)
    q[:, :3] = rng.normal(size=(500, 3)) * 10
    idx, _ = oracle.knn(pts, q, 5)
    _, ref = scipy_spatial.cKDTree(pts[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=5)
    assert (idx == ref).mean() > 0.999                       # float vs double distance can swap near-ties


def test_pcl_get_transformation_is_rz_ry_rx(oracle):
    from multi_sensor_slam_tookit_b200.synth import rot_zyx
    rng = np.random.default_rng(7)
    for _ in range(50):
        pose = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-50, 50, 3)]).astype(np.float32)
        t = oracle.pose_affine(pose)
        assert np.allclose(t[:, :3], rot_zyx(*pose[:3].astype(np.float64)), atol=1e-6)
        assert np.array_equal(t[:, 3], pose[3:])
