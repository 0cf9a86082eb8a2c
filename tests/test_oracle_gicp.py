"""CPU pins of the GICP oracle (oracle/o_gicp.cpp) against independent numpy / scipy restatements of the Open3D
algorithms Multi_LiCa calls (Calibration.py:314-340). Open3D itself is not installable here, so this is what stands
between the oracle and "parity unpinned" for row a17: every building block is checked against a library routine that
was not written for this repository (numpy.linalg, scipy.linalg.sqrtm, scipy.spatial.cKDTree).
Also covers the host-side sharding logic with a world_size-2 gloo run."""
import os
import sys

import numpy as np
import pytest

import gicp_cases as G


def test_voxel_down_sample_against_numpy(oracle):
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.uniform(-30, 30, (5000, 3)), rng.normal(0, 0.02, (2000, 3))])
    for voxel in (0.05, 0.4, 3.0):
        out, rank = oracle.o3d_voxel_down_sample(pts, voxel)
        vmin = pts.min(0) - voxel * 0.5
        ijk = np.floor((pts - vmin) / voxel).astype(np.int64)
        order = np.lexsort((ijk[:, 0], ijk[:, 1], ijk[:, 2]))               # ascending (z, y, x)
        uniq, inv = np.unique(ijk[:, ::-1], axis=0, return_inverse=True)    # rows sorted as (z, y, x)
        inv = inv.reshape(-1)
        assert len(out) == len(uniq)
        assert np.array_equal(rank, inv)
        sums = np.zeros((len(uniq), 3)); np.add.at(sums, inv, pts)
        cnt = np.bincount(inv, minlength=len(uniq))[:, None]
        assert np.allclose(out, sums / cnt, rtol=0, atol=1e-12)
        assert order is not None


def test_normals_against_numpy_eigh(oracle):
    from scipy.spatial import cKDTree
    pts = G.lidar_cloud(0, n_rings=16, n_cols=256)
    nrm, cov = oracle.gicp_normals_covs(pts, 30, 0.005)
    tree = cKDTree(pts)
    _, idx = tree.query(pts, k=30)
    bad = 0
    for i in range(0, len(pts), 7):
        nb = pts[idx[i]]
        C = np.cov(nb.T, bias=True)
        w, V = np.linalg.eigh(C)
        if w[1] - w[0] < 1e-6 * max(w[2], 1e-12):
            continue                                   # direction undefined
        if abs(abs(V[:, 0] @ nrm[i]) - 1.0) > 1e-7:
            bad += 1
    assert bad == 0
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-12)
    # covariance = R diag(eps,1,1) R^T = I - (1 - eps) n n^T away from Open3D's special case
    ok = nrm[:, 0] >= -0.99
    ref = np.eye(3)[None] - (1 - 0.005) * nrm[:, :, None] * nrm[:, None, :]
    assert np.abs(cov[ok] - ref[ok]).max() < 1e-12
    if (~ok).any():
        assert np.abs(cov[~ok] - np.diag([0.005, 1, 1])).max() == 0.0


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def test_linearize_against_scipy_sqrtm(oracle):
    """Open3D's formulas written out with scipy: W = sqrtm(inv(Ct + R Cs R^T)), r = W d, J = W [-[vs]x | I]."""
    from scipy.linalg import sqrtm
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(2)
    tgt = G.lidar_cloud(0, n_rings=16, n_cols=128)
    T_true = G.pose_matrix([0.01, -0.02, 0.03, 0.1, -0.05, 0.02])
    src = (tgt[rng.permutation(len(tgt))[:600]] - T_true[:3, 3]) @ T_true[:3, :3] + rng.normal(0, 0.01, (600, 3))
    _, tc = oracle.gicp_normals_covs(tgt, 30, 0.005)
    _, sc = oracle.gicp_normals_covs(src, 30, 0.005)
    g = oracle.GicpOracle(src, sc, tgt, tc)
    T = G.perturbed(T_true, (0.03, 0.02, -0.01), (0.2, -0.1, 0.3))
    sums, corr = g.linearize(T, 0.5, want_corr=True)
    R, t = T[:3, :3], T[:3, 3]
    vs_all = src @ R.T + t
    d, j = cKDTree(tgt).query(vs_all, k=1)
    hit = d < 0.5
    assert np.array_equal(corr >= 0, hit) and np.array_equal(corr[hit], j[hit])
    JtJ = np.zeros((6, 6)); Jtr = np.zeros(6); r2 = 0.0
    for i in np.nonzero(hit)[0]:
        vs = vs_all[i]
        M = tc[j[i]] + R @ sc[i] @ R.T
        W = np.real(sqrtm(np.linalg.inv(M)))
        J = W @ np.hstack([-skew(vs), np.eye(3)])
        r = W @ (vs - tgt[j[i]])
        JtJ += J.T @ J; Jtr += J.T @ r; r2 += r @ r
    full = np.zeros((6, 6)); full[np.triu_indices(6)] = sums[:21]
    full = full + np.triu(full, 1).T
    assert np.abs(full - JtJ).max() <= 1e-9 * np.abs(JtJ).max()
    assert np.abs(sums[21:27] - Jtr).max() <= 1e-9 * np.abs(Jtr).max()
    assert sums[27] == hit.sum()
    assert abs(sums[28] - (d[hit] ** 2).sum()) <= 1e-9 * sums[28]
    assert abs(sums[29] - r2) <= 1e-9 * r2
    # the update is Open3D's: x = solve(JtJ, -Jtr), T = [Rz(x2) Ry(x1) Rx(x0) | x3..5]
    U, ok = g.solve_update(sums)
    x = np.linalg.solve(JtJ, -Jtr)
    from multi_sensor_slam_tookit_b200.synth import rot_zyx
    assert ok and np.abs(U[:3, :3] - rot_zyx(x[0], x[1], x[2])).max() < 1e-9 and np.abs(U[:3, 3] - x[3:]).max() < 1e-9


def test_register_recovers_rig_transform(oracle):
    src, tgt = G.lidar_cloud(1), G.lidar_cloud(0)
    src, _ = oracle.o3d_voxel_down_sample(src, 0.2)
    tgt, _ = oracle.o3d_voxel_down_sample(tgt, 0.2)
    _, sc = oracle.gicp_normals_covs(src, 30, 0.005)
    _, tc = oracle.gicp_normals_covs(tgt, 30, 0.005)
    truth = G.pair_truth(1, 0)
    res = oracle.GicpOracle(src, sc, tgt, tc).register(G.perturbed(truth), 1.0, 1e-7, 1e-7, 100)
    dT = np.linalg.inv(truth) @ res["transformation"]
    assert np.linalg.norm(dT[:3, 3]) < 0.05 and G.rot_angle(dT[:3, :3]) < np.deg2rad(0.3)
    assert 0 < res["iterations"] < 100 and 0.3 < res["fitness"] <= 1.0


def _gloo_rank(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import pyoracle as O
    from multi_sensor_slam_tookit_b200.gicp import shard_blocks, pair_owner
    import gicp_cases as GC
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    tgt = GC.lidar_cloud(0, n_rings=16, n_cols=128)
    src = tgt[rng.permutation(len(tgt))[:1001]] + rng.normal(0, 0.01, (1001, 3))
    _, tc = O.gicp_normals_covs(tgt, 30, 0.005, threads=2)
    _, sc = O.gicp_normals_covs(src, 30, 0.005, threads=2)
    g = O.GicpOracle(src, sc, tgt, tc, threads=2)
    T = GC.perturbed(np.eye(4), (0.02, 0.01, 0.0), (0.1, 0.0, 0.2))
    blocks = shard_blocks(len(src), rank, world, block=128)
    local = sum(g.linearize(T, 1.0, begin=b, end=e) for b, e in blocks)
    t = torch.from_numpy(local.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                 # the one exchange step of the sharded registration
    full = g.linearize(T, 1.0)
    owners = [pair_owner(i, world) for i in range(5)]
    q.put((rank, blocks[:2], float(np.abs(t.numpy() - full).max() / np.abs(full).max()), float(t[27]), float(full[27]), owners))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_sums():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gloo_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, e0, n0, f0, o0), (r1, s1, e1, n1, f1, o1) = res
    assert s0 == [(0, 128), (256, 384)] and s1 == [(128, 256), (384, 512)]      # blocks dealt round-robin
    assert e0 <= 1e-12 and e1 <= 1e-12 and n0 == f0 == n1 == f1
    assert o0 == [0, 1, 0, 1, 0]


def test_shard_helpers():
    from multi_sensor_slam_tookit_b200.gicp import shard_blocks, shard_size
    for n in (0, 1, 7, 4096, 4097, 1_000_003):
        for world in (1, 2, 3, 8):
            cover = np.zeros(n, np.int32)
            for r in range(world):
                for b, e in shard_blocks(n, r, world):
                    cover[b:e] += 1
                assert shard_size(n, r, world) == sum(e - b for b, e in shard_blocks(n, r, world))
            assert np.all(cover == 1)
            sizes = [shard_size(n, r, world) for r in range(world)]
            assert max(sizes) - min(sizes) <= 4096


def test_icp_error_oracle_against_scipy(oracle):
    """oracle/o_icp.cpp (SensorsCalibration yaw grid search): the error sum against scipy's kd-tree, and the search loop's
    bookkeeping (37 evaluations, the winner is a grid point, the returned transform is GetDeltaT(best_yaw) * init)."""
    from scipy.spatial import cKDTree
    tgt = G.lidar_cloud(0, n_rings=16, n_cols=256).astype(np.float32)
    src = G.lidar_cloud(1, n_rings=16, n_cols=256).astype(np.float32)
    truth = G.pair_truth(1, 0)
    o = oracle.IcpErrorOracle(tgt, src)
    for T in (truth, np.eye(4)):
        q = (src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        d, _ = cKDTree(tgt.astype(np.float64)).query(q.astype(np.float64), k=1)
        ref = float((d ** 2).sum())
        assert abs(o.evaluate(T) - ref) <= 1e-5 * ref
    init = G.perturbed(truth, (0, 0, 0), (0, 0, 0.4))
    r = o.yaw_search(init)
    assert r["evaluations"] == 37 and r["min_error"] <= o.evaluate(init)
    a = np.float32(r["best_yaw"]).astype(np.float64) * np.pi / 180.0
    D = np.eye(4); D[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    assert np.abs(r["transform"] - D @ init).max() < 1e-12
