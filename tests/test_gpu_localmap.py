"""GPU parity of the local-map assembly (mapOptmization.cpp:899-938 extractCloud + the key-frame containers)
against the CPU oracle's transformPointCloud and VoxelGrid, through the C ABI. Everything here is bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_keyframes(c1, n_keys=12, seed=5):
    from multi_sensor_slam_tookit_b200 import synth
    return synth.keyframes_from_map(c1["map_corner"], c1["map_surf"], n_keys, seed)


def oracle_extract(oracle, keys, order, corner_leaf, surf_leaf, poses=None):
    cat_c = np.concatenate([oracle.transform_cloud(keys[k][0], keys[k][2] if poses is None else poses[k]) for k in order]) if order else np.zeros((0, 4), np.float32)
    cat_s = np.concatenate([oracle.transform_cloud(keys[k][1], keys[k][2] if poses is None else poses[k]) for k in order]) if order else np.zeros((0, 4), np.float32)
    ds_c = oracle.voxel_grid(cat_c, corner_leaf)["out"] if len(cat_c) else cat_c
    ds_s = oracle.voxel_grid(cat_s, surf_leaf)["out"] if len(cat_s) else cat_s
    return cat_c, cat_s, ds_c, ds_s


def test_extract_cloud_bit_exact(b2, oracle, c1):
    from multi_sensor_slam_tookit_b200.registration import LocalMap
    keys = make_keyframes(c1)
    lm = LocalMap(0.2, 0.4, surroundingKeyframeSearchRadius=1e9)
    for c, s, p in keys:
        lm.saveKeyFrame(c, s, p)
    for order in ([3, 1, 4, 0, 5, 9, 2, 6], list(range(12)), [7], [2, 2, 5]):
        nc, ns = lm.extractCloud(order)
        cat_c, cat_s, ds_c, ds_s = oracle_extract(oracle, keys, order, 0.2, 0.4)
        assert np.array_equal(lm.get("corner"), cat_c) and np.array_equal(lm.get("surf"), cat_s)      # transform + concatenation order
        assert (nc, ns) == (len(ds_c), len(ds_s))
        assert np.array_equal(lm.get("cornerDS"), ds_c) and np.array_equal(lm.get("surfDS"), ds_s)    # VoxelGrid, centroid sums in input order
    assert lm.extractCloud([]) == (0, 0) and len(lm.get("cornerDS")) == 0


def test_distance_gate_and_cache_semantics(b2, oracle, c1):
    from multi_sensor_slam_tookit_b200.registration import LocalMap
    keys = make_keyframes(c1, n_keys=6)
    lm = LocalMap(0.2, 0.4, surroundingKeyframeSearchRadius=25.0)
    for c, s, p in keys:
        lm.saveKeyFrame(c, s, p)
    last = keys[-1][2][3:6]
    near = [k for k in range(6) if np.sqrt(np.float32(np.sum((keys[k][2][3:6] - last) ** 2, dtype=np.float32))) <= 25.0]
    assert 0 < len(near) < 6
    lm.extractCloud(range(6))
    _, _, ds_c, ds_s = oracle_extract(oracle, keys, near, 0.2, 0.4)
    assert np.array_equal(lm.get("cornerDS"), ds_c) and np.array_equal(lm.get("surfDS"), ds_s)
    # a corrected pose is invisible until the container is cleared (laserCloudMapContainer semantics, :910-922, :1591)
    k = near[0]
    new_pose = keys[k][2].copy(); new_pose[3] += 0.5; new_pose[2] += 0.02
    lm.correctPose(k, new_pose)
    moved = np.sqrt(np.float32(np.sum((new_pose[3:6] - last) ** 2, dtype=np.float32))) <= 25.0 or k == 5
    lm.extractCloud(near)
    assert np.array_equal(lm.get("surfDS"), ds_s)
    lm.clearMapContainer()
    lm.extractCloud(near)
    poses = {i: keys[i][2] for i in range(6)}; poses[k] = new_pose
    _, _, ds_c2, ds_s2 = oracle_extract(oracle, keys, near if moved else [i for i in near if i != k], 0.2, 0.4, poses)
    assert np.array_equal(lm.get("cornerDS"), ds_c2) and np.array_equal(lm.get("surfDS"), ds_s2)
    assert not np.array_equal(ds_s2, ds_s)
    ms, cached = lm.lastGpuMs()
    assert cached == len(near) and ms > 0


def test_scan_to_map_on_the_assembled_map(b2, oracle, c1):
    """kdtree*FromMap->setInputCloud on the device-resident DS clouds gives the same solve as the host path."""
    from multi_sensor_slam_tookit_b200.registration import LocalMap, ScanToMapOptimizer
    keys = make_keyframes(c1)
    lm = LocalMap(0.2, 0.4, surroundingKeyframeSearchRadius=1e9)
    for c, s, p in keys:
        lm.saveKeyFrame(c, s, p)
    lm.extractCloud(range(12))
    a = ScanToMapOptimizer(); a.setInputMapFromLocalMap(lm); a.setInputScan(c1["scan_corner"], c1["scan_surf"])
    b = ScanToMapOptimizer(); b.setInputMap(lm.get("cornerDS"), lm.get("surfDS")); b.setInputScan(c1["scan_corner"], c1["scan_surf"])
    a.transformTobeMapped = c1["pose_guess"].copy(); b.transformTobeMapped = c1["pose_guess"].copy()
    ra = a.scan2MapOptimization(30, record_history=True); rb = b.scan2MapOptimization(30, record_history=True)
    assert ra["iters"] == rb["iters"] and ra["converged"] == rb["converged"]
    assert np.array_equal(ra["pose_history"], rb["pose_history"])
    o = oracle.Scan2Map(4)
    o.set_map(lm.get("cornerDS"), lm.get("surfDS")); o.set_scan(c1["scan_corner"], c1["scan_surf"])
    ref = o.solve(c1["pose_guess"])
    assert ra["iters"] == ref["iters"]
    assert np.abs(ra["pose_history"][:, 3:] - ref["pose_hist"][:, 3:]).max() <= 1e-5
    assert np.abs(ra["pose_history"][:, :3] - ref["pose_hist"][:, :3]).max() <= 1e-6


def test_empty_keyframe_and_bad_index(b2, c1):
    from multi_sensor_slam_tookit_b200.registration import LocalMap
    from multi_sensor_slam_tookit_b200 import capi
    lm = LocalMap()
    lm.saveKeyFrame(np.zeros((0, 4), np.float32), c1["map_surf"][:100], np.zeros(6, np.float32))
    assert lm.extractCloud([0])[0] == 0 and len(lm.get("surf")) == 100
    with pytest.raises(capi.B2Error):
        idx = np.array([3], np.int32)
        capi.check(capi.lib().b2_localmap_extract(lm._h, capi.ptr(idx), 1, None, None))
