"""GPU parity of the nearest-neighbour registration error and the yaw grid search (SURVEY.md §8f N2:
SensorsCalibration lidar2lidar auto_calib/src/registration_icp.cpp:49-100) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import gicp_cases as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clouds():
    tgt = G.lidar_cloud(0, n_rings=32, n_cols=512).astype(np.float32)
    src = G.lidar_cloud(1, n_rings=32, n_cols=512).astype(np.float32)
    tgt = tgt[tgt[:, 2] > -1.5]; src = src[src[:, 2] > -1.5]         # "non-ground" clouds, as the reference feeds
    return tgt, src, G.pair_truth(1, 0)


def test_error_matches_oracle(b2, oracle, clouds):
    from multi_sensor_slam_tookit_b200.registration import ICPRegistrator
    tgt, src, truth = clouds
    r = ICPRegistrator(); r.SetTargetCloud(tgt); r.SetSourceCloud(src)
    o = oracle.IcpErrorOracle(tgt, src)
    for T in (truth, np.eye(4), G.perturbed(truth, (0.3, -0.2, 0.1), (1.0, 2.0, -8.0))):
        e, ref = r.CalculateICPError(T), o.evaluate(T)
        assert abs(e - ref) <= 1e-6 * ref                          # float distances in FLANN, double here
    assert r.CalculateICPError(truth) < 0.2 * r.CalculateICPError(np.eye(4))
    assert r.CalculateICPError(truth) == r.CalculateICPError(truth)   # reproducible: partial sums added in warp order


def test_yaw_search_matches_oracle(b2, oracle, clouds):
    from multi_sensor_slam_tookit_b200.registration import ICPRegistrator
    tgt, src, truth = clouds
    r = ICPRegistrator(); r.SetTargetCloud(tgt); r.SetSourceCloud(src)
    o = oracle.IcpErrorOracle(tgt, src)
    # the grid is in "degrees of degrees" (GetDeltaT converts radians once more): the reachable yaw range is about +-0.9 deg
    for dyaw_deg in (0.35, -0.6):
        init = G.perturbed(truth, (0.0, 0.0, 0.0), (0.0, 0.0, dyaw_deg))
        res, ref = r.RegistrationByICP(init), o.yaw_search(init)
        assert res["evaluations"] == ref["evaluations"] == 37
        assert res["best_yaw"] == ref["best_yaw"]                   # same grid point wins
        assert np.array_equal(res["transform"], ref["transform"])
        assert abs(res["min_error"] - ref["min_error"]) <= 1e-6 * ref["min_error"]
        assert res["min_error"] <= r.CalculateICPError(init) and res["gpu_ms"] > 0


def test_edge_cases(b2, clouds):
    from multi_sensor_slam_tookit_b200.registration import ICPRegistrator
    from multi_sensor_slam_tookit_b200 import capi
    tgt, src, truth = clouds
    r = ICPRegistrator()
    with pytest.raises(capi.B2Error):
        r.CalculateICPError(np.eye(4))
    r.SetTargetCloud(tgt); r.SetSourceCloud(np.zeros((0, 3), np.float32))
    assert r.CalculateICPError(np.eye(4)) == 0.0
    r.SetSourceCloud(src[:1]); r.SetTargetCloud(tgt[:1])
    d = src[0].astype(np.float64) - tgt[0].astype(np.float64)
    assert abs(r.CalculateICPError(np.eye(4)) - float(d @ d)) <= 1e-6 * float(d @ d)
    # PointXYZI stride (32 B) accepted as is
    p32 = np.zeros((len(tgt), 8), np.float32); p32[:, :3] = tgt
    r.SetTargetCloud(p32); r.SetSourceCloud(src)
    r2 = ICPRegistrator(); r2.SetTargetCloud(tgt); r2.SetSourceCloud(src)
    assert r.CalculateICPError(truth) == r2.CalculateICPError(truth)
