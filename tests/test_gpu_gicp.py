"""GPU parity of the Multi_LiCa GICP path against the CPU oracle, through the C ABI.

Bars: voxel membership, neighbour sets and correspondence sets bit-exact (ties by index);
J^T J / J^T r / rmse <= 1e-9 relative; final transform <= 1e-5 m / 1e-6 rad; same iteration count.
"""
import numpy as np
import pytest

import gicp_cases as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gicp(b2):
    from multi_sensor_slam_tookit_b200 import gicp as mod
    return mod


@pytest.fixture(scope="module")
def pair(gicp, oracle):
    """Lidar 1 -> lidar 0 of the rig, voxel 0.1, normals on both sides; device clouds and oracle arrays."""
    src_raw, tgt_raw = G.lidar_cloud(1), G.lidar_cloud(0)
    src = gicp.PointCloud(src_raw).voxel_down_sample(0.1)
    tgt = gicp.PointCloud(tgt_raw).voxel_down_sample(0.1)
    src.estimate_normals(); tgt.estimate_normals()
    eps = 0.005
    sp, tp = src.points, tgt.points
    sn, sc = oracle.gicp_normals_covs(sp, 30, eps)
    tn, tc = oracle.gicp_normals_covs(tp, 30, eps)
    return dict(src=src, tgt=tgt, sp=sp, tp=tp, sn=sn, tn=tn, sc=sc, tc=tc, eps=eps,
                truth=G.pair_truth(1, 0), oracle=oracle.GicpOracle(sp, sc, tp, tc))


def test_voxel_down_sample_bit_exact(gicp, oracle):
    rng = np.random.default_rng(3)
    cases = [G.lidar_cloud(2), rng.uniform(-20, 20, (30000, 3)), rng.normal(0, 0.01, (500, 3)) + [1e4, -1e4, 50.0]]
    for pts in cases:
        for voxel in (0.05, 0.5, 7.3):
            out, rank = gicp.PointCloud(pts).voxel_down_sample(voxel, return_voxel_rank=True)
            ref, ref_rank = oracle.o3d_voxel_down_sample(pts, voxel)
            assert len(out) == len(ref)
            assert np.array_equal(rank, ref_rank)
            assert np.array_equal(out.points, ref)          # same sums in the same order: bit-exact means


def test_voxel_down_sample_edge_cases(gicp, oracle):
    assert len(gicp.PointCloud(np.zeros((0, 3))).voxel_down_sample(0.1)) == 0
    one = gicp.PointCloud(np.array([[1.0, 2.0, 3.0]])).voxel_down_sample(0.1)
    assert np.array_equal(one.points, [[1.0, 2.0, 3.0]])
    same = np.tile([[0.3, -0.2, 9.0]], (1000, 1))
    out = gicp.PointCloud(same).voxel_down_sample(0.05)
    assert len(out) == 1 and np.allclose(out.points, same[:1], atol=1e-12)
    with pytest.raises(RuntimeError):
        gicp.PointCloud(same).voxel_down_sample(0.0)
    # 64-bit voxel keys (more than 2^32 voxels in the bounding box): the two-word sort path
    rng = np.random.default_rng(5)
    wide = rng.uniform(-400, 400, (20000, 3))
    out, rank = gicp.PointCloud(wide).voxel_down_sample(0.3, return_voxel_rank=True)
    ref, ref_rank = oracle.o3d_voxel_down_sample(wide, 0.3)
    assert np.array_equal(rank, ref_rank) and np.array_equal(out.points, ref)
    # float32 input is widened exactly
    f32 = wide.astype(np.float32)
    assert np.array_equal(gicp.PointCloud(f32).points, f32.astype(np.float64))


def test_estimate_normals_matches_oracle(pair, gicp, oracle):
    for dev, ref in ((pair["src"], pair["sn"]), (pair["tgt"], pair["tn"])):
        nrm = dev.normals
        assert nrm.shape == ref.shape
        exact = np.all(nrm == ref, axis=1).mean()
        # same neighbours, same summation order, same Jacobi: bit-exact up to libm-free arithmetic
        assert exact > 0.999, exact
        assert np.abs(nrm - ref).max() < 1e-9
    # fewer than 3 points -> (0, 0, 1)
    tiny = gicp.PointCloud(np.array([[0.0, 0, 0], [1.0, 0, 0]]))
    tiny.estimate_normals()
    assert np.array_equal(tiny.normals, [[0, 0, 1.0], [0, 0, 1.0]])
    # knn larger than the cloud: all points are neighbours
    rng = np.random.default_rng(1)
    few = rng.uniform(-1, 1, (12, 3)); few[:, 2] *= 0.01
    c = gicp.PointCloud(few); c.estimate_normals()
    ref, _ = oracle.gicp_normals_covs(few, 30, 0.005)
    assert np.abs(c.normals - ref).max() < 1e-12


def test_estimate_normals_uneven_density(gicp, oracle):
    """Raw ring scan (no downsampling): neighbour radius varies by two orders of magnitude, rings must widen."""
    pts = G.lidar_cloud(3, n_rings=16, n_cols=256)
    c = gicp.PointCloud(pts); c.estimate_normals()
    ref, _ = oracle.gicp_normals_covs(pts, 30, 0.005)
    assert np.all(c.normals == ref, axis=1).mean() > 0.999
    assert np.abs(c.normals - ref).max() < 1e-9


def test_estimate_normals_lattice_ties_and_far_coordinates(gicp, oracle):
    """The neighbour heap orders by the fp32 image of the squared distance and settles equal images exactly. A
    lattice is nothing but equal distances (the 30th neighbour is one of many at the same radius: the smaller index wins, as in
    the oracle), and 6 km from the origin distinct fp64 distances share fp32 images all the time."""
    gx, gy = np.meshgrid(np.arange(60) * 0.25, np.arange(50) * 0.25, indexing="ij")
    z = 0.125 * ((np.arange(60)[:, None] + 2 * np.arange(50)[None, :]) % 3)          # three interleaved height levels
    lattice = np.stack([gx.ravel(), gy.ravel(), z.ravel()], 1)
    rng = np.random.default_rng(5)
    lattice = lattice[rng.permutation(len(lattice))]                                    # index order unrelated to position
    for shift in (np.zeros(3), np.array([6000.0, -2500.0, 40.0])):
        pts = lattice + shift
        c = gicp.PointCloud(pts); c.estimate_normals()
        ref, _ = oracle.gicp_normals_covs(pts, 30, 0.005)
        assert np.all(c.normals == ref, axis=1).mean() > 0.999
        assert np.abs(c.normals - ref).max() < 1e-9


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check_sums(sums, ref):
    assert sums[27] == ref[27]                                         # n_corr exact
    JtJ, JtJ_ref = sums[:21], ref[:21]
    assert rel_err(JtJ, JtJ_ref) <= 1e-9
    assert rel_err(sums[21:27], ref[21:27]) <= 1e-9
    assert abs(sums[28] - ref[28]) <= 1e-9 * max(ref[28], 1e-300)
    assert abs(sums[29] - ref[29]) <= 1e-9 * max(ref[29], 1e-300)


def test_linearize_matches_oracle(pair, gicp):
    g = gicp.GeneralizedICP(1.0, pair["eps"])
    g.setInputTarget(pair["tgt"]); g.setInputSource(pair["src"])
    for T in (G.perturbed(pair["truth"]), pair["truth"], np.eye(4)):
        sums, corr = g.linearize(T, want_correspondences=True)
        ref, ref_corr = pair["oracle"].linearize(T, 1.0, want_corr=True)
        assert np.array_equal(corr, ref_corr)                          # correspondence set bit-exact
        check_sums(sums, ref)
    info = g.indexInfo()
    assert info["target_cell_edge"] > 0 and info["shard_points"] == len(pair["sp"])


def test_linearize_far_apart_and_small_radius(pair, gicp):
    """Neighbours beyond the 3x3x3 block (shell expansion) and a radius smaller than a cell."""
    g = gicp.GeneralizedICP(6.0, pair["eps"])
    g.setInputTarget(pair["tgt"]); g.setInputSource(pair["src"])
    T = pair["truth"].copy(); T[:3, 3] += [2.5, -1.5, 0.7]
    sums, corr = g.linearize(T, want_correspondences=True)
    ref, ref_corr = pair["oracle"].linearize(T, 6.0, want_corr=True)
    assert np.array_equal(corr, ref_corr)
    check_sums(sums, ref)
    g.setParams(max_correspondence_distance=0.05)
    sums, corr = g.linearize(pair["truth"], want_correspondences=True)
    ref, ref_corr = pair["oracle"].linearize(pair["truth"], 0.05, want_corr=True)
    assert np.array_equal(corr, ref_corr)
    check_sums(sums, ref)
    # no overlap at all
    T[:3, 3] += 1e4
    sums, corr = g.linearize(T, want_correspondences=True)
    assert sums[27] == 0 and np.all(corr == -1) and np.all(sums == 0)


def test_coarse_pass_screening_far_coordinates_and_ties(pair, gicp, oracle):
    """The coarse pass decides on an fp32 copy of the target relative to the cell corners and on boxes quantised to 1/256 cell
    before it touches the fp64 points. Clouds 7 km from the origin (where a plain fp32 coordinate would be off by
    half a millimetre), duplicated target points (exact ties: the smaller index wins) and an offset that sends nearly every query
    through the coarse pass: the correspondence set still equals the oracle's bit for bit."""
    rng = np.random.default_rng(20261018)
    sel_s = rng.choice(len(pair["sp"]), 6000, replace=False); sel_t = rng.choice(len(pair["tp"]), 9000, replace=False)
    shift = np.array([5000.0, -4800.0, 130.0])
    sp = pair["sp"][sel_s] + shift
    tp = pair["tp"][sel_t] + shift
    dup = rng.choice(len(tp), 600, replace=False)
    tp = np.concatenate([tp, tp[dup]])                                  # ties between equal points
    sn, tn = pair["sn"][sel_s], np.concatenate([pair["tn"][sel_t], pair["tn"][sel_t][dup]])
    src, tgt = gicp.PointCloud(sp), gicp.PointCloud(tp)
    src.normals = sn; tgt.normals = tn
    sc, tc = pair["sc"][sel_s], np.concatenate([pair["tc"][sel_t], pair["tc"][sel_t][dup]])     # covariances do not move with the shift
    ref_o = oracle.GicpOracle(sp, sc, tp, tc)
    for radius, off in ((1.0, [0.45, -0.3, 0.25]), (2.5, [1.2, 0.8, -0.4]), (0.6, [0.0, 0.0, 0.0])):
        g = gicp.GeneralizedICP(radius, pair["eps"])
        g.setInputTarget(tgt); g.setInputSource(src)
        # the pose moves the source about its own position (a rotation about the far origin would throw it kilometres away)
        T = np.eye(4); T[:3, :3] = G.perturbed(np.eye(4))[:3, :3]
        c0 = sp.mean(0); T[:3, 3] = c0 - T[:3, :3] @ c0 + np.array(off)
        sums, corr = g.linearize(T, want_correspondences=True)
        ref, ref_corr = ref_o.linearize(T, radius, want_corr=True)
        assert np.array_equal(corr, ref_corr)
        assert (corr >= 0).sum() > 100
        check_sums(sums, ref)


def test_align_matches_oracle(pair, gicp):
    init = G.perturbed(pair["truth"])
    res = gicp.registration_generalized_icp(pair["src"], pair["tgt"], 1.0, init,
                                            gicp.TransformationEstimationForGeneralizedICP(pair["eps"]),
                                            gicp.ICPConvergenceCriteria(1e-7, 1e-7, 100))
    ref = pair["oracle"].register(init, 1.0, 1e-7, 1e-7, 100)
    assert res.iterations == ref["iterations"]
    assert res.fitness == ref["fitness"]
    assert abs(res.inlier_rmse - ref["inlier_rmse"]) <= 1e-9
    dT = np.linalg.inv(ref["transformation"]) @ res.transformation
    assert np.linalg.norm(dT[:3, 3]) <= 1e-5 and G.rot_angle(dT[:3, :3]) <= 1e-6
    # and it actually calibrates: within 3 cm / 0.2 deg of the rig truth
    dT = np.linalg.inv(pair["truth"]) @ res.transformation
    assert np.linalg.norm(dT[:3, 3]) < 0.03 and G.rot_angle(dT[:3, :3]) < np.deg2rad(0.2)
    assert len(res.correspondence_set) == round(res.fitness * len(pair["sp"]))
    assert res.gpu_launches >= res.iterations + 1 and len(res.fitness_history) == res.iterations + 1


def test_align_iteration_cap_and_zero_iterations(pair, gicp):
    init = G.perturbed(pair["truth"])
    for cap in (0, 1, 3):
        res = gicp.registration_generalized_icp(pair["src"], pair["tgt"], 1.0, init,
                                                gicp.TransformationEstimationForGeneralizedICP(pair["eps"]),
                                                gicp.ICPConvergenceCriteria(1e-7, 1e-7, cap))
        ref = pair["oracle"].register(init, 1.0, 1e-7, 1e-7, cap)
        assert res.iterations == ref["iterations"] == cap
        assert res.fitness == ref["fitness"] and abs(res.inlier_rmse - ref["inlier_rmse"]) <= 1e-9
        assert np.abs(res.transformation - ref["transformation"]).max() <= 1e-9


def test_sharded_sums_add_up(pair, gicp):
    """The source shards of SURVEY.md §8e (C5): blocks of 4096 cell-sorted points dealt round-robin. Per-rank sums add up to
    the single-GPU sums (no communicator needed for single evaluations), the shards are disjoint and cover the source."""
    g = gicp.GeneralizedICP(1.0, pair["eps"])
    g.setInputTarget(pair["tgt"]); g.setInputSource(pair["src"])
    T = G.perturbed(pair["truth"])
    full, full_corr = g.linearize(T, want_correspondences=True)
    n = len(pair["sp"])
    from multi_sensor_slam_tookit_b200 import capi
    for world in (2, 3, 8):
        total = np.zeros(30); seen = np.zeros(n, bool); sizes = []
        for rank in range(world):
            capi.check(capi.lib().b2_gicp_set_shard(g._h, rank, world, None))
            s, c = g.linearize(T, want_correspondences=True)
            total += s
            info = g.indexInfo()
            sizes.append(info["shard_points"])
            assert info["shard_block_points"] == gicp.SHARD_BLOCK_POINTS
            assert info["shard_points"] == gicp.shard_size(n, rank, world)
            mine = c >= 0
            assert not np.any(seen & mine)
            assert np.array_equal(c[mine], full_corr[mine])
            seen |= mine
        assert sum(sizes) == n
        assert np.array_equal(seen, full_corr >= 0)
        assert total[27] == full[27] and rel_err(total[:27], full[:27]) <= 1e-12
    capi.check(capi.lib().b2_gicp_set_shard(g._h, 0, 1, None))


def test_errors(gicp):
    g = gicp.GeneralizedICP()
    with pytest.raises(Exception):
        g.align(np.eye(4))                                   # no clouds yet
    c = gicp.PointCloud(np.random.default_rng(0).uniform(0, 1, (100, 3)))
    with pytest.raises(Exception):
        g.setInputTarget(c)                                  # no normals
    with pytest.raises(Exception):
        c.estimate_normals(knn=64)                           # beyond the supported neighbourhood


def test_two_gpu_sharded_align(pair, gicp, tmp_path):
    """C5 in miniature on 2 GPUs (skipped on a 1-GPU box): NCCL all-reduce of the 30 sums per iteration."""
    from multi_sensor_slam_tookit_b200 import capi
    if capi.lib().b2_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, os, json
    from conftest import ROOT
    np.savez(tmp_path / "in.npz", sp=pair["sp"], tp=pair["tp"], sn=pair["sn"], tn=pair["tn"], init=G.perturbed(pair["truth"]))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tests", "_gicp_rank.py"), str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r0 = json.load(open(tmp_path / "rank0.json")); r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["T"] == r1["T"] and r0["iterations"] == r1["iterations"]         # every rank solves the same system
    single = gicp.GeneralizedICP(1.0, pair["eps"])
    s = gicp.PointCloud(pair["sp"]); s.normals = pair["sn"]
    t = gicp.PointCloud(pair["tp"]); t.normals = pair["tn"]
    single.setInputTarget(t); single.setInputSource(s)
    ref = single.align(G.perturbed(pair["truth"]))
    assert r0["iterations"] == ref.iterations and r0["fitness"] == ref.fitness
    assert np.abs(np.array(r0["T"]) - ref.transformation).max() <= 1e-9
    # sharded set-up: all-gathered points and normals bit-identical to the single-GPU calls on every rank, and the registration
    # on row slices of the source (normals from estimate_normals on both sides, so compared with a single-GPU run of the same)
    assert r0["pts_equal"] and r1["pts_equal"] and r0["nrm_equal"] and r1["nrm_equal"]
    assert r0["T2"] == r1["T2"] and r0["slice"][1] == r1["slice"][0]
    s1 = gicp.PointCloud(pair["sp"]); t1 = gicp.PointCloud(pair["tp"])
    s1.estimate_normals(); t1.estimate_normals()
    one = gicp.GeneralizedICP(1.0, pair["eps"]); one.setInputTarget(t1); one.setInputSource(s1)
    ref2 = one.align(G.perturbed(pair["truth"]))
    assert r0["iterations2"] == ref2.iterations and abs(r0["fitness2"] - ref2.fitness) <= 1e-12
    assert np.abs(np.array(r0["T2"]) - ref2.transformation).max() <= 1e-9
    assert r0["T5"] == r1["T5"] and r0["iterations5"] == ref2.iterations and abs(r0["fitness5"] - ref2.fitness) <= 1e-12
    assert np.abs(np.array(r0["T5"]) - ref2.transformation).max() <= 1e-9           # Morton-block shards: same registration
    # fused linearise + exchange (CUDA IPC peer stores): bit-identical on both ranks, and to the NCCL path (two ranks: a + b either way)
    assert r0["fused"] and r1["fused"], "the two GPUs of the box cannot map each other's memory"
    assert r0["T3"] == r1["T3"] and r0["T3"] == r0["T2"] and r0["iterations3"] == r0["iterations2"] and r0["T4"] == r0["T3"]
    # b2_gicp_exchange_setup (handles over the communicator) on the Morton-block shards: the same bits as their NCCL run
    assert r0["fused6"] and r1["fused6"]
    assert r0["T6"] == r1["T6"] and r0["T6"] == r0["T5"] and r0["iterations6"] == r0["iterations5"]
