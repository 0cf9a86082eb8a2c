"""The reference's own real clouds through the path.

Data: tests/golden/real_lidar2lidar_0001.npz = Multi_LiCa/data/demo/lidar_{1,2,3}.pcd (a 64-ring roof lidar, 92 677 returns,
and two tilted side lidars, 8 572 / 9 248 returns), start = auto_calib/data/0001/initial_extrinsic.txt (yaw +-90 deg, roll =
pitch = 0 although the side lidars are pitched by ~45 deg).

What is pinned here that is not ours:
* the reference publishes no extrinsics for these three clouds, so there is no ground truth to hit. What it publishes is the
  accuracy of the Multi_LiCa pipeline, `evaluation/config.yaml:4-13`: <= 4.24 cm and <= 0.345 deg per axis between calibration
  and ground truth. The CPU tests assert that the two *independent algorithms* of this path, restated independently
  (Open3D GICP with Multi_LiCa's parameters, PCL NDT with multi_lidar_calibrator's), land inside that envelope of each other
  on both real pairs; that GICP reaches the same transform from three starts up to 2 deg / 20 cm apart; and that the
  alignment is real by a measure none of our code computes: the share of source returns within 10 cm of a target return
  (scipy cKDTree) goes from ~0.1 % at the shipped start to ~20 %.
* the GPU tests run exactly the same calls through the C ABI and must reproduce the oracle on this data (same iteration
  counts, correspondence sets bit-exact, transforms <= 1e-5 m / 1e-6 rad), plus the front end on the real 64-ring sweep.
"""
import numpy as np
import pytest

import real_cases as RC

ENV_M, ENV_DEG = 0.0424, 0.3454          # evaluation/config.yaml:4-13, max per-axis |calibration - ground truth|


def test_envelope_constants_are_the_published_table():
    fx = RC.fixture()
    assert abs(fx["envelope"][0] - 0.042378) < 1e-6 and abs(fx["envelope"][1] - 0.345341) < 1e-6
    assert fx["envelope"][0] <= ENV_M and fx["envelope"][1] <= ENV_DEG
    assert fx["published_abs_error"].shape == (3, 6)


@pytest.fixture(scope="module")
def oracle_runs(oracle):
    """Oracle GICP (Calibration.py:292-345 flow) and NDT (multi_lidar_calibrator.cpp:28-121 flow) on both real pairs."""
    fx = RC.fixture()
    tgt_raw = fx["lidar_1"].astype(np.float64)
    tgt, _ = oracle.o3d_voxel_down_sample(tgt_raw, RC.GICP["voxel_size"])
    _, tcov = oracle.gicp_normals_covs(tgt, 30, RC.GICP["epsilon"])
    runs = {"tgt": tgt, "tcov": tcov}
    for name in ("lidar_2", "lidar_3"):
        src_raw = fx[name].astype(np.float64)
        src, _ = oracle.o3d_voxel_down_sample(src_raw, RC.GICP["voxel_size"])
        _, scov = oracle.gicp_normals_covs(src, 30, RC.GICP["epsilon"])
        g = oracle.GicpOracle(src, scov, tgt, tcov)
        reg = lambda init, g=g: g.register(init, RC.GICP["max_corresp_dist"], RC.GICP["rel_fitness"], RC.GICP["rel_rmse"], RC.GICP["max_iterations"])
        init = RC.initial_guess(name)
        r0 = reg(init)
        nd = oracle.NdtOracle(RC.NDT["resolution"], RC.NDT["step_size"], RC.NDT["epsilon"], RC.NDT["max_iterations"])
        nd.set_target(fx["lidar_1"])
        child = oracle.voxel_grid(RC.xyzi(fx[name]), RC.NDT["voxel_size"])["out"][:, :3]
        nd.set_source(child)
        runs[name] = dict(src=src, scov=scov, gicp=g, reg=reg, init=init, r0=r0, ndt=nd, child=child, raw=fx[name].astype(np.float64))
    return runs


@pytest.mark.parametrize("name", ["lidar_2", "lidar_3"])
def test_oracle_gicp_on_real_pair_is_a_real_alignment(oracle_runs, name):
    from scipy.spatial import cKDTree
    r = oracle_runs[name]
    T = r["r0"]["transformation"]
    assert 10 <= r["r0"]["iterations"] < RC.GICP["max_iterations"] and r["r0"]["fitness"] > 0.7
    tree = cKDTree(RC.fixture()["lidar_1"].astype(np.float64))
    share = lambda M: float((tree.query(r["raw"] @ M[:3, :3].T + M[:3, 3])[0] < 0.1).mean())
    before, after = share(r["init"]), share(T)
    assert before < 0.005 and after > 0.15, (before, after)
    # the side lidars are tilted: the shipped start has pitch 0, the registration finds ~45 deg, yaw stays near +-90
    e = RC.euler_deg(T)
    assert 40 < e[1] < 50 and abs(abs(e[2]) - 90) < 5


@pytest.mark.parametrize("name", ["lidar_2", "lidar_3"])
def test_oracle_gicp_same_answer_from_perturbed_starts(oracle_runs, name):
    r = oracle_runs[name]
    T0 = r["r0"]["transformation"]
    for xyz, rpy in (((0.05, 0.05, -0.03), (0.5, -0.5, 1.0)), ((-0.1, 0.08, 0.05), (-1.0, 0.7, -1.5)), ((0.2, -0.15, 0.1), (2.0, 1.5, -2.0))):
        r1 = r["reg"](RC.perturb(r["init"], xyz, rpy))
        dm, dd = RC.per_axis_difference(T0, r1["transformation"])
        assert dm.max() <= 0.1 * ENV_M and dd.max() <= 0.1 * ENV_DEG, (xyz, rpy, dm, dd)


@pytest.mark.parametrize("name", ["lidar_2", "lidar_3"])
def test_oracle_ndt_and_gicp_agree_within_published_envelope(oracle_runs, name):
    """Two different objective functions (NDT's voxel Gaussians on the raw target vs GICP's plane-to-plane residuals on the
    0.05 m clouds), restated separately, on real data: NDT started at the GICP result, and at two disturbed copies of it,
    stays or returns inside the accuracy the reference publishes for this pipeline."""
    r = oracle_runs[name]
    Tg = r["r0"]["transformation"]
    for xyz, rpy in (((0, 0, 0), (0, 0, 0)), ((0.2, -0.2, 0.1), (2.0, 2.0, -2.0)), ((0.1, 0.1, -0.1), (1.0, -1.0, 1.0))):
        a = r["ndt"].align(RC.perturb(Tg, xyz, rpy).astype(np.float32))
        assert a["converged"]
        Tn = np.asarray(a["transformation"], np.float64).reshape(4, 4)
        dm, dd = RC.per_axis_difference(Tg, Tn)
        assert dm.max() <= ENV_M and dd.max() <= ENV_DEG, (name, xyz, rpy, dm, dd)


# ================================================================================================ GPU, through the C ABI
@pytest.fixture(scope="module")
def gpu_clouds(b2):
    from multi_sensor_slam_tookit_b200 import gicp
    fx = RC.fixture()
    out = {}
    for name in ("lidar_1", "lidar_2", "lidar_3"):
        c = gicp.PointCloud(fx[name].astype(np.float64)).voxel_down_sample(RC.GICP["voxel_size"])
        c.estimate_normals()
        out[name] = c
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lidar_2", "lidar_3"])
def test_gpu_gicp_on_real_pair_equals_oracle(b2, oracle, oracle_runs, gpu_clouds, name):
    from multi_sensor_slam_tookit_b200 import gicp
    r = oracle_runs[name]
    src, tgt = gpu_clouds[name], gpu_clouds["lidar_1"]
    assert np.array_equal(src.points, r["src"]) and np.array_equal(tgt.points, oracle_runs["tgt"])        # voxel means bit-exact
    res = gicp.registration_generalized_icp(
        src, tgt, RC.GICP["max_corresp_dist"], r["init"], gicp.TransformationEstimationForGeneralizedICP(RC.GICP["epsilon"]),
        gicp.ICPConvergenceCriteria(RC.GICP["rel_fitness"], RC.GICP["rel_rmse"], RC.GICP["max_iterations"]))
    ref = r["r0"]
    assert res.iterations == ref["iterations"]
    D = np.linalg.inv(ref["transformation"]) @ res.transformation
    ang = np.arccos(np.clip((np.trace(D[:3, :3]) - 1) / 2, -1, 1))
    assert np.linalg.norm(D[:3, 3]) <= 1e-5 and ang <= 1e-6, (np.linalg.norm(D[:3, 3]), ang)
    assert abs(res.fitness - ref["fitness"]) <= 1e-12 and abs(res.inlier_rmse - ref["inlier_rmse"]) <= 1e-9
    # correspondence sets at the start and at the end, bit-exact
    g = gicp.GeneralizedICP(RC.GICP["max_corresp_dist"], RC.GICP["epsilon"])
    g.setInputTarget(tgt); g.setInputSource(src)
    for T in (r["init"], ref["transformation"]):
        sums, corr = g.linearize(T, want_correspondences=True)
        rs, rcorr = r["gicp"].linearize(T, RC.GICP["max_corresp_dist"], want_corr=True)
        assert np.array_equal(corr, rcorr)
        assert np.abs(sums[:27] - rs[:27]).max() <= 1e-9 * np.abs(rs[:27]).max()
    # and the GPU result sits inside the published envelope of the (independent) NDT answer
    a = r["ndt"].align(ref["transformation"].astype(np.float32))
    dm, dd = RC.per_axis_difference(res.transformation, np.asarray(a["transformation"], np.float64).reshape(4, 4))
    assert dm.max() <= ENV_M and dd.max() <= ENV_DEG


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lidar_2", "lidar_3"])
def test_gpu_ndt_on_real_pair_equals_oracle(b2, oracle, oracle_runs, name):
    from multi_sensor_slam_tookit_b200 import ndt, registration
    fx = RC.fixture()
    r = oracle_runs[name]
    vg = registration.VoxelGrid(); vg.setLeafSize(*(RC.NDT["voxel_size"],) * 3); vg.setInputCloud(RC.xyzi(fx[name]))
    child = vg.filter()[:, :3]
    assert np.array_equal(child, r["child"])                                   # PCL VoxelGrid on real data, bit-exact
    n = ndt.NormalDistributionsTransform()
    n.setTransformationEpsilon(RC.NDT["epsilon"]); n.setStepSize(RC.NDT["step_size"]); n.setResolution(RC.NDT["resolution"])
    n.setMaximumIterations(RC.NDT["max_iterations"])
    n.setInputTarget(fx["lidar_1"]); n.setInputSource(np.ascontiguousarray(child))
    for guess in (RC.perturb(r["r0"]["transformation"], (0.2, -0.2, 0.1), (2.0, 2.0, -2.0)), r["init"]):
        guess = guess.astype(np.float32)
        n.align(guess)
        ref = r["ndt"].align(guess)
        assert n.getFinalNumIteration() == ref["iterations"] and n.hasConverged() == ref["converged"]
        D = np.linalg.inv(np.asarray(ref["transformation"], np.float64).reshape(4, 4)) @ n.getFinalTransformation().astype(np.float64)
        ang = np.arccos(np.clip((np.trace(D[:3, :3]) - 1) / 2, -1, 1))
        assert np.linalg.norm(D[:3, 3]) <= 1e-5 and ang <= 1e-6 + 4e-4 * 0
        assert n.lastGpuMs()["evaluations"] == ref["evaluations"]


@pytest.mark.gpu
def test_gpu_front_end_on_real_64_ring_sweep(b2, oracle):
    """imageProjection + featureExtraction on the real roof-lidar sweep (N_SCAN 64, Horizon_SCAN 1800, ring and time fields
    as shipped), with a gyro table through the a3 helper: the same bars as the synthetic C2 tests."""
    from multi_sensor_slam_tookit_b200 import frontend
    raw = RC.lidar1_xyzirt()
    t0 = 1000.0
    stamp = np.arange(t0 - 0.05, t0 + 0.2, 1.0 / 400.0)
    gyro = np.stack([0.2 * np.sin(9 * stamp), 0.15 * np.cos(4 * stamp), 0.5 + 0 * stamp], 1)
    info = frontend.imu_deskew_info((stamp, None, gyro), t0, t0 + 0.1)
    ref_info = oracle.imu_deskew_info(stamp, None, gyro, t0, t0 + 0.1)
    assert info["imuAvailable"] and all(np.array_equal(a, b) for a, b in zip(info["imu"], ref_info["imu"]))
    fe = frontend.ScanFrontEnd(N_SCAN=64, Horizon_SCAN=1800)
    got = fe.projectPointCloud(raw, imu=info["imu"], timeScanCur=t0, want_images=True)
    ref = oracle.project(raw, 64, 1800, imu=ref_info["imu"], t_cur=t0)
    assert len(got["extracted"]) == len(ref["extracted"]) > 80000
    assert np.array_equal(got["pointColInd"], ref["pointColInd"]) and np.array_equal(got["pointRange"], ref["pointRange"])
    assert np.array_equal(got["startRingIndex"], ref["startRingIndex"]) and np.array_equal(got["endRingIndex"], ref["endRingIndex"])
    assert np.array_equal(got["range_mat"], ref["range_mat"])
    assert np.abs(got["extracted"][:, :3] - ref["extracted"][:, :3]).max() <= 2e-5 * max(1.0, float(np.abs(ref["extracted"][:, :3]).max()) / 10)
    # features from the oracle's projection on both sides (so the comparison is bit-level, as in test_gpu_frontend.py)
    fe2 = frontend.ScanFrontEnd(N_SCAN=64, Horizon_SCAN=1800)
    g2 = fe2.projectPointCloud(raw, imu=None, timeScanCur=t0, deskew=False)
    r2 = oracle.project(raw, 64, 1800, imu=None, t_cur=t0, deskew=False)
    assert np.array_equal(g2["extracted"], r2["extracted"])
    gf = fe2.extractFeatures(want_arrays=True)
    rf = oracle.extract_features(r2)
    assert np.array_equal(gf["label"], rf["label"]) and np.array_equal(gf["corner"], rf["corner"]) and np.array_equal(gf["surf"], rf["surf"])
    assert len(gf["corner"]) > 1000 and len(gf["surf"]) > 5000
