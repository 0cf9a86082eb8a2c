"""GPU parity of the multi-lidar fusion front end (PointClouds_Fusion fusion_pointclouds.cpp:55-115) against the
numpy restatement in oracle/pyoracle.py, through the C ABI. Bit-exact: transforms, order, both box filters."""
import numpy as np
import pytest

import gicp_cases as G

pytestmark = pytest.mark.gpu


def rig_clouds():
    clouds, T = [], []
    for i in range(4):
        p = G.lidar_cloud(i, n_rings=16, n_cols=256).astype(np.float32)
        c = np.zeros((len(p), 4), np.float32); c[:, :3] = p; c[:, 3] = np.arange(len(p)) % 255
        clouds.append(c)
        T.append(None if i == 0 else G.pair_truth(i, 0))           # children into the parent frame (:62-73)
    # the reference's order: pc_local_1 + pc_trans_4 + pc_trans_3 + pc_trans_2 (:80-89)
    order = [0, 3, 2, 1]
    return [clouds[k] for k in order], [T[k] for k in order]


def test_fusion_bit_exact(b2, oracle):
    from multi_sensor_slam_tookit_b200.registration import FusionPc
    clouds, T = rig_clouds()
    clouds[1][5, 0] = np.nan; clouds[2][7, 2] = np.inf              # non-finite points: dropped by the pass-through only
    f = FusionPc()
    ext = ((-40.0, -35.5, -2.0), (45.25, 30.0, 8.0)); inn = ((-2.0, -1.5, -3.0), (2.5, 1.5, 0.5))
    for e, i in ((None, None), (ext, None), (None, inn), (ext, inn)):
        out = f.fuse(clouds, T, e, i)
        ref = oracle.fuse_clouds(clouds, T, e, i)
        assert out.shape == ref.shape
        assert np.array_equal(out, ref, equal_nan=True)
        assert f.n_fused == sum(len(c) for c in clouds)
    assert 0 < len(f.fuse(clouds, T, ext, inn)) < f.n_fused and f.lastGpuMs() > 0
    # two clouds, one transformed (lidar_fusion.cpp:239-252, 333-334), PointXYZI stride
    wide = [np.zeros((len(c), 8), np.float32) for c in clouds[:2]]
    for w, c in zip(wide, clouds[:2]):
        w[:, :3] = c[:, :3]; w[:, 4] = c[:, 3]
    out = f.fuse([wide[1], wide[0]], [T[1], None])
    ref = oracle.fuse_clouds([clouds[1], clouds[0]], [T[1], None])
    assert np.array_equal(out, ref, equal_nan=True)
    assert len(f.fuse([], [])) == 0
