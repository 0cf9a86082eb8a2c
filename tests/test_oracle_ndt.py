"""CPU pins of the NDT oracle (oracle/o_ndt.cpp) against independent numpy restatements of PCL's published algorithm
(PCL itself is not installable here): voxel statistics, the SVD solve, the analytic gradient / Hessian against finite
differences of the score over fixed (point, voxel) pairs, and the end-to-end calibration of the synthetic rig."""
import numpy as np

import gicp_cases as G
import ndt_cases as N


def test_voxel_statistics_against_numpy(oracle):
    tgt, _, _ = N.pair(n_rings=32, n_cols=256)
    o = oracle.NdtOracle(1.0)
    nv = o.set_target(tgt)
    v = o.voxels()
    assert nv == len(v["index"]) > 50
    inv = np.float32(1.0) / np.float32(1.0)
    ijk = (np.floor(tgt * inv) - v["min_b"].astype(np.float32)).astype(np.int64)
    lin = ijk[:, 0] + ijk[:, 1] * v["div_b"][0] + ijk[:, 2] * v["div_b"][0] * v["div_b"][1]
    uniq, counts = np.unique(lin, return_counts=True)
    assert np.array_equal(uniq[counts >= 6], v["index"])
    for k in range(0, nv, 7):
        P = tgt[lin == v["index"][k]].astype(np.float64)
        n = len(P)
        assert n == v["npts"][k] or v["npts"][k] == -1
        assert np.allclose(P.mean(0), v["mean"][k], rtol=0, atol=1e-12)
        assert np.allclose(tgt[lin == v["index"][k]].mean(0), v["centroid"][k], atol=1e-4)
        if v["npts"][k] < 0:
            continue
        C = np.cov(P.T, bias=True) * (n - 1.0) / n           # PCL's single-pass covariance, scaled by (n-1)/n as upstream
        w, V = np.linalg.eigh(C)
        w = np.maximum(w, 0.01 * w[2]) if w[0] < 0.01 * w[2] else w
        Cn = V @ np.diag(w) @ V.T
        ref = np.linalg.inv(Cn)
        assert np.abs(v["icov"][k] - ref).max() <= 1e-6 * np.abs(ref).max()


def test_svd_solve_against_numpy(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        A = rng.normal(size=(6, 6)); H = A + A.T + rng.normal() * np.eye(6)
        b = rng.normal(size=6)
        x = oracle.svd_solve6(H, b)
        assert np.abs(x - np.linalg.solve(H, b)).max() <= 1e-9 * max(1.0, np.abs(x).max())
    # rank-deficient: the minimum-norm solution, as JacobiSVD::solve returns
    u = rng.normal(size=(6, 3)); H = u @ u.T; b = H @ rng.normal(size=6)
    assert np.abs(oracle.svd_solve6(H, b) - np.linalg.lstsq(H, b, rcond=1e-12)[0]).max() < 1e-8


def _rot(p):
    cx, sx, cy, sy, cz, sz = np.cos(p[3]), np.sin(p[3]), np.cos(p[4]), np.sin(p[4]), np.cos(p[5]), np.sin(p[5])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def test_derivatives_against_finite_differences(oracle):
    """score(p) = sum over FIXED (point, voxel) pairs of -d1 exp(-d2/2 (T(p)x - mu)^T S^-1 (T(p)x - mu)), eq. 6.9-6.10; its numeric
    gradient and Hessian (double, numpy) must match the oracle's analytic eq. 6.12 / 6.13 at the same pose."""
    tgt, src, truth = N.pair(n_rings=32, n_cols=256, oracle=oracle)
    src = src[::3]
    o = oracle.NdtOracle(1.0)
    o.set_target(tgt); o.set_source(src)
    v = o.voxels()
    p0 = N.pose_vector(truth) + np.array([0.05, -0.03, 0.02, 0.004, -0.003, 0.01])
    s0, g0, H0, pairs = o.derivatives(p0)
    # the pairs the oracle used: float transform, float centroid distance < resolution^2, voxels with >= 6 points
    Tf = oracle.ndt_pose_to_matrix(p0)
    xt = (src @ Tf[:3, :3].T + Tf[:3, 3]).astype(np.float32)
    from scipy.spatial import cKDTree
    tree = cKDTree(v["centroid"].astype(np.float64))
    nb = tree.query_ball_point(xt.astype(np.float64), 1.0 - 1e-6)
    pi = np.repeat(np.arange(len(src)), [len(b) for b in nb]); vi = np.concatenate([np.asarray(b, int) for b in nb])
    assert abs(len(pi) - pairs) <= 2                       # ties on the radius aside, the same pair set
    c1, c2 = 10 * (1 - 0.55), 0.55 / 1.0 ** 3
    d3 = -np.log(c2); d1 = -np.log(c1 + c2) - d3; d2 = -2 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)
    X = src[pi].astype(np.float64); MU = v["mean"][vi]; IC = v["icov"][vi]

    def score(p):
        q = X @ _rot(p).T + p[:3] - MU
        e = np.exp(-d2 / 2 * np.einsum("ni,nij,nj->n", q, IC, q))
        return float((-d1 * e).sum())

    h = 1e-5
    gn = np.array([(score(p0 + h * np.eye(6)[i]) - score(p0 - h * np.eye(6)[i])) / (2 * h) for i in range(6)])
    assert abs(score(p0) - s0) <= 2e-4 * abs(s0)           # float vs double transform of the points
    assert np.abs(gn - g0).max() <= 2e-3 * np.abs(g0).max()
    hh = 1e-4
    Hn = np.zeros((6, 6))
    for i in range(6):
        for j in range(6):
            ei, ej = hh * np.eye(6)[i], hh * np.eye(6)[j]
            Hn[i, j] = (score(p0 + ei + ej) - score(p0 + ei - ej) - score(p0 - ei + ej) + score(p0 - ei - ej)) / (4 * hh * hh)
    assert np.abs(Hn - H0).max() <= 5e-3 * np.abs(H0).max()


def test_align_calibrates_the_rig(oracle):
    tgt, src, truth = N.pair(oracle=oracle)
    o = oracle.NdtOracle(1.0, 0.1, 0.01, 400)
    o.set_target(tgt); o.set_source(src)
    r = o.align(G.perturbed(truth, (0.15, -0.05, 0.03), (1.0, -0.5, 3.0)).astype(np.float32))
    dT = np.linalg.inv(truth) @ r["transformation"].astype(np.float64)
    assert r["converged"] and 3 < r["iterations"] < 60
    assert np.linalg.norm(dT[:3, 3]) < 0.02 and G.rot_angle(dT[:3, :3]) < np.deg2rad(0.1)
    assert r["transformation_probability"] > 1.0 and o.fitness(r["transformation"]) < 1.0
