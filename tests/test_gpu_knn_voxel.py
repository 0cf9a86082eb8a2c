"""GPU parity (through the C ABI) for the grid k-NN index and the VoxelGrid, against the CPU oracle.
Bar: bit-exact indices / voxel assignments / squared distances; centroids bit-exact in the pinned summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cloud(rng, n, scale=(40, 40, 6)):
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(-1, 1, (n, 3)) * np.array(scale)
    p[:, 3] = rng.uniform(0, 100, n)
    return p


def test_knn5_bit_exact_on_c1_maps(b2, oracle, c1):
    from multi_sensor_slam_tookit_b200.registration import KdTreeFLANN
    for m, q in ((c1["map_corner"], c1["scan_corner"]), (c1["map_surf"], c1["scan_surf"])):
        # queries in the map frame: apply the guess pose with the oracle's transform
        qs = oracle.transform_cloud(q, c1["pose_guess"])
        kd = KdTreeFLANN(1.0)
        kd.setInputCloud(m)
        idx, d2 = kd.nearestKSearch(qs, 5)
        oi, od = oracle.knn(m, qs, 5)
        inside = od < 1.0                                   # exactness is promised for neighbours closer than max_dist
        assert inside[:, 4].mean() > 0.5
        assert np.array_equal(idx[inside], oi[inside])
        assert np.array_equal(d2[inside], od[inside])
        assert np.all(idx[~inside] == -1) and np.all(np.isinf(d2[~inside]))


def test_knn_ties_duplicates_and_strides(b2, oracle):
    from multi_sensor_slam_tookit_b200.registration import KdTreeFLANN
    rng = np.random.default_rng(21)
    pts = _cloud(rng, 30000, (10, 10, 2))
    pts[:2000, :3] = pts[2000:4000, :3]                     # duplicates -> equal distances, order by index
    q = _cloud(rng, 5000, (11, 11, 2.5))
    q[:500, :3] = pts[:500, :3]
    for k in (1, 5, 8):
        kd = KdTreeFLANN(1.0)
        kd.setInputCloud(pts)
        idx, d2 = kd.nearestKSearch(q, k)
        oi, od = oracle.knn(pts, q, k, brute=True)
        inside = od < 1.0
        assert np.array_equal(idx[inside], oi[inside]) and np.array_equal(d2[inside], od[inside])
    # PCL's 32-byte PointXYZI layout goes in as it is
    wide = np.zeros((len(pts), 8), np.float32); wide[:, :3] = pts[:, :3]; wide[:, 3] = 1; wide[:, 4] = pts[:, 3]
    kd = KdTreeFLANN(1.0); kd.setInputCloud(wide)
    idx2, _ = kd.nearestKSearch(q, 5)
    kd = KdTreeFLANN(1.0); kd.setInputCloud(pts)
    idx1, _ = kd.nearestKSearch(q, 5)
    assert np.array_equal(idx1, idx2)


def test_knn_edge_cases(b2):
    from multi_sensor_slam_tookit_b200.registration import KdTreeFLANN
    kd = KdTreeFLANN(1.0)
    kd.setInputCloud(np.zeros((0, 4), np.float32))
    idx, d2 = kd.nearestKSearch(np.zeros((3, 4), np.float32), 5)
    assert np.all(idx == -1) and np.all(np.isinf(d2))
    kd.setInputCloud(np.array([[0, 0, 0, 0], [0.5, 0, 0, 0], [np.nan, 0, 0, 0]], np.float32))
    idx, d2 = kd.nearestKSearch(np.array([[0.1, 0, 0, 0], [100, 100, 100, 0], [np.nan, 0, 0, 0]], np.float32), 5)
    assert list(idx[0][:2]) == [0, 1] and np.all(idx[0][2:] == -1) and np.all(idx[1:] == -1)


@pytest.mark.parametrize("n,leaf,scale", [(200000, 0.4, (60, 60, 8)), (5000, 0.2, (5, 5, 1)), (100000, 0.1, (8, 8, 3))])
def test_voxel_grid_bit_exact(b2, oracle, n, leaf, scale):
    from multi_sensor_slam_tookit_b200.registration import VoxelGrid
    rng = np.random.default_rng(n)
    pts = _cloud(rng, n, scale)
    vg = VoxelGrid(); vg.setLeafSize(leaf, leaf, leaf); vg.setInputCloud(pts)
    out, vop = vg.filter(return_voxel_index=True)
    ref = oracle.voxel_grid(pts, leaf)
    assert not vg.refused
    assert np.array_equal(vop, ref["voxel_of_point"])        # voxel assignment of every point
    assert out.shape == ref["out"].shape
    assert np.array_equal(out, ref["out"])                   # same order, same float sums


def test_voxel_grid_on_c1_and_layouts(b2, oracle, c1):
    from multi_sensor_slam_tookit_b200.registration import VoxelGrid
    cloud = np.concatenate([c1["map_surf"], c1["map_surf"] + np.float32(0.05)])
    vg = VoxelGrid(); vg.setLeafSize(0.4, 0.4, 0.4); vg.setInputCloud(cloud)
    out = vg.filter()
    assert np.array_equal(out, oracle.voxel_grid(cloud, 0.4)["out"])
    wide = np.zeros((len(cloud), 8), np.float32); wide[:, :3] = cloud[:, :3]; wide[:, 3] = 1; wide[:, 4] = cloud[:, 3]
    vg.setInputCloud(wide)
    outw = vg.filter()
    assert np.array_equal(outw[:, :3], out[:, :3]) and np.array_equal(outw[:, 4], out[:, 3]) and np.all(outw[:, 3] == 1)
    vg.setInputCloud(np.ascontiguousarray(cloud[:, :3]))     # PointXYZ (multi_lidar_calibrator.h:70)
    out3 = vg.filter()
    assert np.array_equal(out3, out[:, :3])


def test_voxel_grid_edge_cases(b2, oracle):
    from multi_sensor_slam_tookit_b200.registration import VoxelGrid
    vg = VoxelGrid(); vg.setLeafSize(0.2, 0.2, 0.2)
    vg.setInputCloud(np.zeros((0, 4), np.float32))
    assert len(vg.filter()) == 0
    one = np.array([[1.5, -2.5, 0.25, 7.0]], np.float32)
    vg.setInputCloud(one)
    assert np.array_equal(vg.filter(), one)
    far = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2]], np.float32)
    vg.setLeafSize(0.001, 0.001, 0.001); vg.setInputCloud(far)
    assert np.array_equal(vg.filter(), far) and vg.refused
    pts = np.array([[0.01, 0.01, 0.01, 1], [0.02, 0.02, 0.02, 3], [5, 5, 5, 1]], np.float32)
    vg.setLeafSize(0.1, 0.1, 0.1); vg.setMinimumPointsNumberPerVoxel(2); vg.setInputCloud(pts)
    assert np.array_equal(vg.filter(), oracle.voxel_grid(pts, 0.1, min_points=2)["out"])


def test_voxel_grid_idempotent_at_scale(b2):
    # size-independent property at a size the oracle is not asked to match: filtering the output again with the
    # same leaf cannot merge anything when each voxel already holds exactly one centroid inside its own cell
    from multi_sensor_slam_tookit_b200.registration import VoxelGrid
    rng = np.random.default_rng(33)
    pts = _cloud(rng, 4_000_000, (200, 200, 10))
    vg = VoxelGrid(); vg.setLeafSize(0.5, 0.5, 0.5); vg.setInputCloud(pts)
    out, vop = vg.filter(return_voxel_index=True)
    assert len(out) == len(np.unique(vop))
    vg.setInputCloud(out)
    out2 = vg.filter()
    assert len(out2) == len(out)
    # mass is conserved: count-weighted centroid mean == global mean
    cnt = np.bincount(np.unique(vop, return_inverse=True)[1])
    assert np.allclose((out[:, :3].astype(np.float64) * cnt[:, None]).sum(0) / len(pts), pts[:, :3].astype(np.float64).mean(0), atol=1e-3)
