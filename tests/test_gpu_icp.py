"""GPU parity of the loop-closure ICP (pcl::IterativeClosestPoint as mapOptmization.cpp:559-586 configures it)
against the CPU oracle, through the C ABI: same correspondences (float FLANN distances, ties by index) -> same iteration
count and convergence flag, final transformation within 1e-5 m / 1e-6 rad, fitness score within 1e-6 relative."""
import numpy as np
import pytest

import gicp_cases as G

pytestmark = pytest.mark.gpu


def submaps(oracle):
    """cureKeyframeCloud / prevKeyframeCloud in miniature: two scans of the scene from nearby poses, in a common frame up to
    a loop-closure drift, downsampled with VoxelGrid(0.4) as loopFindNearKeyframes does (mappingSurfLeafSize)."""
    a = G.lidar_cloud(0, n_rings=32, n_cols=512).astype(np.float32)
    b = G.lidar_cloud(4, n_rings=32, n_cols=512).astype(np.float32)
    T_ab = G.pair_truth(4, 0)
    b_in_a = (b.astype(np.float64) @ T_ab[:3, :3].T + T_ab[:3, 3]).astype(np.float32)
    drift = G.perturbed(np.eye(4), (0.4, -0.3, 0.1), (0.5, -0.3, 2.0))
    cur = (b_in_a.astype(np.float64) @ drift[:3, :3].T + drift[:3, 3]).astype(np.float32)
    ds = lambda p: oracle.voxel_grid(np.c_[p, np.zeros(len(p), np.float32)], 0.4)["out"][:, :3].copy()     # noqa: E731
    return ds(cur), ds(a), np.linalg.inv(drift)


def run(icp_cls, src, tgt, **kw):
    icp = icp_cls()
    icp.setMaxCorrespondenceDistance(kw.get("d", 30.0)); icp.setMaximumIterations(kw.get("it", 100))
    icp.setTransformationEpsilon(kw.get("te", 1e-6)); icp.setEuclideanFitnessEpsilon(kw.get("fe", 1e-6)); icp.setRANSACIterations(0)
    icp.setInputSource(src); icp.setInputTarget(tgt)
    out = icp.align(want_output=True)
    return icp, out


def test_loop_closure_icp_matches_oracle(b2, oracle):
    from multi_sensor_slam_tookit_b200.registration import IterativeClosestPoint
    src, tgt, truth = submaps(oracle)
    assert len(src) >= 300 and len(tgt) >= 1000                      # the reference's own size gate (:552)
    for kw in (dict(), dict(d=2.0), dict(it=3), dict(te=1e-3, fe=1e-3)):
        icp, out = run(IterativeClosestPoint, src, tgt, **kw)
        ref = oracle.icp_align(src, tgt, kw.get("d", 30.0), kw.get("it", 100), kw.get("te", 1e-6), kw.get("fe", 1e-6))
        assert icp.getFinalNumIteration() == ref["iterations"] and icp.hasConverged() == ref["converged"]
        T, Tr = icp.getFinalTransformation().astype(np.float64), ref["transformation"].astype(np.float64)
        dT = np.linalg.inv(Tr) @ T
        assert np.linalg.norm(dT[:3, 3]) <= 1e-5 and G.rot_angle(dT[:3, :3]) <= 1e-6 + 3e-4 * 0
        fs = icp.getFitnessScore()
        assert abs(fs - ref["fitness_score"]) <= 1e-6 * ref["fitness_score"]
        exp = src @ T[:3, :3].T.astype(np.float32) + T[:3, 3].astype(np.float32)
        assert np.abs(out - exp).max() < 1e-3
    # it closes the loop: the drift is recovered, and the reference's acceptance test (:573) passes
    icp, _ = run(IterativeClosestPoint, src, tgt)
    dT = np.linalg.inv(truth) @ icp.getFinalTransformation().astype(np.float64)
    # (point-to-point ICP between scans taken from two different mounting poses: decimetre-level, not millimetre-level)
    assert np.linalg.norm(dT[:3, 3]) < 0.2 and G.rot_angle(dT[:3, :3]) < np.deg2rad(1.5)
    assert icp.hasConverged() and icp.getFitnessScore() < 1.0          # partial overlap: the far returns dominate the mean


def test_icp_guess_and_degenerate_inputs(b2, oracle):
    from multi_sensor_slam_tookit_b200.registration import IterativeClosestPoint
    from multi_sensor_slam_tookit_b200 import capi
    src, tgt, truth = submaps(oracle)
    # a guess is applied to the working cloud first and ends up in the final transformation
    icp = IterativeClosestPoint(); icp.setMaximumIterations(50); icp.setInputSource(src); icp.setInputTarget(tgt)
    icp.align(guess=truth.astype(np.float32))
    dT = np.linalg.inv(truth) @ icp.getFinalTransformation().astype(np.float64)
    assert np.linalg.norm(dT[:3, 3]) < 0.2
    # fewer than three correspondences: not converged, identity returned (PCL's "Not enough correspondences")
    icp = IterativeClosestPoint(); icp.setMaxCorrespondenceDistance(0.5); icp.setMaximumIterations(20)
    icp.setInputSource(src + np.float32(1000.0)); icp.setInputTarget(tgt); icp.align()
    assert not icp.hasConverged() and icp.getFinalNumIteration() == 0 and np.array_equal(icp.getFinalTransformation(), np.eye(4, dtype=np.float32))
    with pytest.raises(capi.B2Error):
        IterativeClosestPoint().align()
    with pytest.raises(capi.B2Error):
        icp.setRANSACIterations(5)
