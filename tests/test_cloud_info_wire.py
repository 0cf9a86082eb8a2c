"""lio_sam/cloud_info on the wire. CPU part: the host-only parser of the C ABI (b2_cloud_info_parse) against the
struct-based serialiser of oracle/pyoracle.py; GPU part: the device-packed message bytes against the same serialiser, the
featureExtraction and mapOptimization hand-offs against the host-array paths they replace."""
import numpy as np
import pytest

META = dict(seq=7, stamp=(1700000000, 123456789), frame_id="base_link", lidarFrame="lidar_link", imuAvailable=1, odomAvailable=1,
            imuRollInit=0.01, imuPitchInit=-0.02, imuYawInit=1.5, initialGuess=(1.0, -2.0, 0.5, 0.01, 0.02, 0.03))


def _random_info(rng, n_scan=4, horizon=32, m=50, nc=7, ns=19):
    return dict(startRingIndex=rng.integers(0, 100, n_scan), endRingIndex=rng.integers(0, 100, n_scan),
                pointColInd=rng.integers(0, horizon, n_scan * horizon), pointRange=rng.uniform(1, 50, n_scan * horizon).astype(np.float32),
                cloud_deskewed=rng.normal(0, 10, (m, 4)).astype(np.float32), cloud_corner=rng.normal(0, 10, (nc, 4)).astype(np.float32),
                cloud_surface=rng.normal(0, 10, (ns, 4)).astype(np.float32))


def test_parse_against_struct_serialiser(oracle):
    """No GPU: b2_cloud_info_parse is host code. Unaligned frame_id lengths shift every later field."""
    from multi_sensor_slam_tookit_b200.frontend import parse_cloud_info
    from multi_sensor_slam_tookit_b200 import capi
    rng = np.random.default_rng(3)
    for stage in (0, 1):
        for frame in ("", "a", "base_link", "odom1"):
            info = _random_info(rng, m=int(rng.integers(0, 60)))
            meta = dict(META, frame_id=frame, lidarFrame=frame + "x")
            msg = oracle.serialize_cloud_info(stage, **meta, **info)
            v = parse_cloud_info(msg)
            assert v["seq"] == 7 and v["stamp"] == META["stamp"] and v["frame_id"] == frame
            assert v["imuAvailable"] == 1 and v["odomAvailable"] == 1
            assert np.float32(v["imuYawInit"]) == np.float32(1.5) and np.allclose(v["initialGuess"], META["initialGuess"], rtol=1e-7)
            if stage == 0:
                for k in ("startRingIndex", "endRingIndex", "pointColInd", "pointRange"):
                    assert np.array_equal(v[k], info[k])
                assert v["cloud_corner"]["width"] == 0 and v["cloud_corner"]["n_fields"] == 0
            else:
                assert all(len(v[k]) == 0 for k in ("startRingIndex", "endRingIndex", "pointColInd", "pointRange"))
                for k in ("cloud_corner", "cloud_surface"):
                    assert np.array_equal(v[k]["points"][:, [0, 1, 2, 4]], info[k]) and np.all(v[k]["points"][:, 3] == 1.0)
            d = v["cloud_deskewed"]
            assert (d["width"], d["height"], d["point_step"], d["offsets"], d["is_dense"]) == (len(info["cloud_deskewed"]), 1, 32, (0, 4, 8, 16), 1)
            assert np.array_equal(d["points"][:, [0, 1, 2, 4]], info["cloud_deskewed"])
            assert v["key_frame_map"]["width"] == 0
            # truncated and padded messages are refused, not read past
            raw = np.frombuffer(msg, np.uint8)
            view = capi.CloudInfoView()
            import ctypes as C
            for cut in (0, 3, 17, len(msg) // 2, len(msg) - 1):
                part = np.ascontiguousarray(raw[:cut])
                assert capi.lib().b2_cloud_info_parse(capi.ptr(part) if cut else capi.ptr(raw), cut, C.byref(view)) == -1      # B2_ERR_ARG
            longer = np.concatenate([raw, np.zeros(1, np.uint8)])
            assert capi.lib().b2_cloud_info_parse(capi.ptr(longer), longer.size, C.byref(view)) == -1      # B2_ERR_ARG


def _scan(seed):
    from multi_sensor_slam_tookit_b200 import synth
    return synth.ring_scan(synth.CityBlock(), (0.0, 0.0, 0.3, 0.0, -24.0, 1.8), n_rings=16, n_cols=1800, elev_deg=(-15.0, 15.0), seed=seed)


@pytest.mark.gpu
def test_stage_messages_bit_exact_and_round_trip(b2, oracle):
    from multi_sensor_slam_tookit_b200.frontend import ScanFrontEnd, parse_cloud_info
    fe = ScanFrontEnd(16, 1800)
    stale_col = np.zeros(16 * 1800, np.int32); stale_rng = np.zeros(16 * 1800, np.float32)
    for seed, frame in ((5, "base_link"), (6, "odom")):             # second scan: entries past its count keep the first scan's values
        raw = _scan(seed)
        g = fe.projectPointCloud(raw)
        m = len(g["extracted"])
        stale_col[:m] = g["pointColInd"]; stale_rng[:m] = g["pointRange"]
        meta = dict(META, frame_id=frame)
        msg0 = fe.publishClouds(**meta)
        ref0 = oracle.serialize_cloud_info(0, **meta, startRingIndex=g["startRingIndex"], endRingIndex=g["endRingIndex"],
                                           pointColInd=stale_col, pointRange=stale_rng, cloud_deskewed=g["extracted"])
        assert msg0.tobytes() == ref0
        f = fe.extractFeatures(want_arrays=True)
        msg1 = fe.publishFeatureCloud(**meta)
        ref1 = oracle.serialize_cloud_info(1, **meta, cloud_deskewed=g["extracted"], cloud_corner=f["corner"], cloud_surface=f["surf"])
        assert msg1.tobytes() == ref1
        # featureExtraction in another process: load the stage-0 message into a fresh handle, same features
        fe2 = ScanFrontEnd(16, 1800)
        f2 = fe2.laserCloudInfoHandler(msg0)
        assert fe2.n_extracted == m
        assert np.array_equal(f2["corner"], f["corner"]) and np.array_equal(f2["corner_idx"], f["corner_idx"]) and np.array_equal(f2["surf"], f["surf"])
        assert fe2.publishFeatureCloud(**meta).tobytes() == ref1
        v = parse_cloud_info(msg1)
        assert np.array_equal(v["cloud_corner"]["points"][:, [0, 1, 2, 4]], f["corner"])
    # wrong geometry and wrong stage are refused
    from multi_sensor_slam_tookit_b200 import capi
    with pytest.raises(capi.B2Error):
        ScanFrontEnd(32, 1800).laserCloudInfoHandler(msg0)
    with pytest.raises(capi.B2Error):
        ScanFrontEnd(16, 1800).publishClouds()
    fe3 = ScanFrontEnd(16, 1800); fe3.projectPointCloud(raw)
    with pytest.raises(capi.B2Error):
        fe3.publishFeatureCloud()


@pytest.mark.gpu
def test_map_optimization_hand_offs(b2, oracle):
    """laserCloudInfoHandler + downsampleCurrentScan: message -> device VoxelGrids -> optimiser, against the host-array path."""
    from multi_sensor_slam_tookit_b200.frontend import ScanFrontEnd
    from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer, VoxelGrid
    d = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "c1_input.npz"))
    fe = ScanFrontEnd(16, 1800)
    fe.projectPointCloud(_scan(5))
    f = fe.extractFeatures()
    msg1 = fe.publishFeatureCloud(**META)
    dsc, dss = VoxelGrid(), VoxelGrid()
    dsc.setLeafSize(0.2, 0.2, 0.2); dss.setLeafSize(0.4, 0.4, 0.4)          # mappingCornerLeafSize / mappingSurfLeafSize
    ref_c = oracle.voxel_grid(f["corner"], 0.2)["out"]; ref_s = oracle.voxel_grid(f["surf"], 0.4)["out"]
    poses = []
    for how in ("host", "message", "device"):
        g = ScanToMapOptimizer()
        g.setInputMap(d["map_corner"], d["map_surf"])
        if how == "host":
            dsc.setInputCloud(f["corner"]); c = dsc.filter(); dss.setInputCloud(f["surf"]); s = dss.filter()
            g.setInputScan(c, s)
        elif how == "message":
            info = g.laserCloudInfoHandler(msg1, dsc, dss)
            assert info["imuAvailable"] == 1
        else:
            g.setInputScanFromFrontEnd(fe, dsc, dss)
        gc, gs = g.getInputScan()
        assert np.array_equal(gc, ref_c) and np.array_equal(gs, ref_s)
        if how != "host":
            assert (g.laserCloudCornerLastDSNum, g.laserCloudSurfLastDSNum) == (len(ref_c), len(ref_s))
        g.transformTobeMapped = d["pose_guess"].copy()
        r = g.scan2MapOptimization(30, want_matP=False)
        poses.append((np.array(g.transformTobeMapped).copy(), r["iters"]))
    for p, it in poses[1:]:
        assert np.array_equal(p, poses[0][0]) and it == poses[0][1]
    # empty feature clouds go through (laserCloudCornerLastDSNum = 0)
    g = ScanToMapOptimizer(); g.setInputMap(d["map_corner"], d["map_surf"])
    g.setInputScanDownsampled(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), dsc, dss)
    assert g.laserCloudCornerLastDSNum == 0 and g.laserCloudSurfLastDSNum == 0
