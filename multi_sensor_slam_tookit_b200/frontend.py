"""Host-side mirror of LIO-SAM's per-scan front end on top of the C ABI (b2_scan_*).

  ScanFrontEnd.projectPointCloud  <- ImageProjection::projectPointCloud + deskewPoint + cloudExtraction
                                     (liosam_ws/src/LIO-SAM/src/imageProjection.cpp:446-598)
  ScanFrontEnd.extractFeatures    <- FeatureExtraction::calculateSmoothness + markOccludedPoints + extractFeatures
                                     (liosam_ws/src/LIO-SAM/src/featureExtraction.cpp:81-238)
  ScanFrontEnd.publishClouds / publishFeatureCloud / laserCloudInfoHandler  <- the lio_sam/cloud_info hand-offs between the
                                     stages (imageProjection.cpp:600-605, featureExtraction.cpp:66-79, 240-258), ROS1 wire bytes
  parse_cloud_info                <- host-only view of a serialised message
Outputs carry the names of msg/cloud_info.msg. Everything computes on the GPU; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import capi


class ScanFrontEnd:
    def __init__(self, N_SCAN=16, Horizon_SCAN=1800, downsampleRate=1, lidarMinRange=1.0, lidarMaxRange=1000.0,
                 edgeThreshold=1.0, surfThreshold=0.1, odometrySurfLeafSize=0.4):
        L = capi.lib()
        p = capi.ScanParams()
        L.b2_scan_default_params(C.byref(p))
        p.n_scan, p.horizon_scan, p.downsample_rate = N_SCAN, Horizon_SCAN, downsampleRate
        p.lidar_min_range, p.lidar_max_range = lidarMinRange, lidarMaxRange
        p.edge_threshold, p.surf_threshold, p.odometry_surf_leaf_size = edgeThreshold, surfThreshold, odometrySurfLeafSize
        self.params = p
        self._h = C.c_void_p()
        capi.check(L.b2_scan_create(C.byref(self._h), C.byref(p)))
        self.n_extracted = 0

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_scan_destroy(self._h)
        except Exception:
            pass

    def projectPointCloud(self, laserCloudIn, imu=None, timeScanCur=0.0, deskew=True, want_images=False):
        """laserCloudIn: PointXYZIRT records (n*32 bytes). imu: (imuTime, imuRotX, imuRotY, imuRotZ) float64 or None."""
        raw = np.ascontiguousarray(laserCloudIn).view(np.uint8).reshape(-1)
        n = raw.size // 32
        N, H = self.params.n_scan, self.params.horizon_scan
        cells = N * H
        if imu is None:
            it = irx = iry = irz = None
            n_imu = 0
        else:
            it, irx, iry, irz = (np.ascontiguousarray(a, np.float64) for a in imu)
            n_imu = len(it)
        ext = np.empty((cells, 4), np.float32)
        col = np.empty(cells, np.int32)
        rng = np.empty(cells, np.float32)
        sr = np.empty(N, np.int32)
        er = np.empty(N, np.int32)
        rm = np.empty(cells, np.float32) if want_images else None
        fc = np.empty((cells, 4), np.float32) if want_images else None
        m = C.c_size_t(0)
        capi.check(capi.lib().b2_scan_project(self._h, capi.ptr(raw), n, capi.ptr(it), capi.ptr(irx), capi.ptr(iry), capi.ptr(irz), n_imu,
                                              float(timeScanCur), 1 if deskew else -1, C.byref(m), capi.ptr(ext), capi.ptr(col),
                                              capi.ptr(rng), capi.ptr(sr), capi.ptr(er), capi.ptr(rm), capi.ptr(fc)))
        self.n_extracted = m.value
        out = dict(extracted=ext[:m.value].copy(), pointColInd=col[:m.value].copy(), pointRange=rng[:m.value].copy(),
                   startRingIndex=sr, endRingIndex=er)
        if want_images:
            out["range_mat"] = rm.reshape(N, H)
            out["full_cloud"] = fc
        return out

    def extractFeatures(self, want_arrays=False):
        N, H = self.params.n_scan, self.params.horizon_scan
        M = self.n_extracted
        corner = np.empty((N * 120, 4), np.float32)
        cidx = np.empty(N * 120, np.int32)
        surf = np.empty((N * H, 4), np.float32)
        curv = np.empty(max(M, 1), np.float32) if want_arrays else None
        picked = np.empty(max(M, 1), np.int32) if want_arrays else None
        label = np.empty(max(M, 1), np.int32) if want_arrays else None
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib().b2_scan_extract_features(self._h, C.byref(nc), capi.ptr(corner), capi.ptr(cidx), C.byref(ns), capi.ptr(surf),
                                                       capi.ptr(curv), capi.ptr(picked), capi.ptr(label)))
        out = dict(corner=corner[:nc.value].copy(), corner_idx=cidx[:nc.value].copy(), surf=surf[:ns.value].copy())
        if want_arrays:
            out.update(curvature=curv[:M], picked_mask=picked[:M], label=label[:M])
        return out

    def lastGpuMs(self):
        ms = C.c_float(0)
        capi.check(capi.lib().b2_scan_last_gpu_ms(self._h, C.byref(ms)))
        return ms.value

    def imuDeskewInfo(self, imuQueue, timeScanCur, timeScanEnd):
        """ImageProjection::imuDeskewInfo (imageProjection.cpp:305-362), see imu_deskew_info below; the returned dict's
        'imu' entry is what projectPointCloud takes."""
        return imu_deskew_info(imuQueue, timeScanCur, timeScanEnd)

    # ---- lio_sam/cloud_info between the stages
    def _write(self, stage, meta):
        m = cloud_info_meta(**(meta or {}))
        n = C.c_size_t(0)
        capi.check(capi.lib().b2_scan_write_cloud_info(self._h, C.byref(m), stage, None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint8)
        capi.check(capi.lib().b2_scan_write_cloud_info(self._h, C.byref(m), stage, capi.ptr(out), out.size, C.byref(n)))
        return out

    def publishClouds(self, **meta):
        """ImageProjection::publishClouds (imageProjection.cpp:600-605): the deskew/cloud_info message bytes."""
        return self._write(0, meta)

    def publishFeatureCloud(self, **meta):
        """FeatureExtraction::publishFeatureCloud (featureExtraction.cpp:248-258): the feature/cloud_info message bytes."""
        return self._write(1, meta)

    def laserCloudInfoHandler(self, msg):
        """FeatureExtraction::laserCloudInfoHandler (:66-79) up to extractFeatures: load a deskew/cloud_info message."""
        raw = np.ascontiguousarray(np.frombuffer(msg, np.uint8))
        m = C.c_size_t(0)
        capi.check(capi.lib().b2_scan_set_from_cloud_info(self._h, capi.ptr(raw), raw.size, C.byref(m)))
        self.n_extracted = m.value
        return self.extractFeatures()


QUEUE_LENGTH = 2000          # imageProjection.cpp:45


def imu_deskew_info(imuQueue, timeScanCur, timeScanEnd, capacity=QUEUE_LENGTH):
    """imuQueue: (stamp (n,), orientation_xyzw (n, 4) or None, angular_velocity (n, 3)) in arrival order, after imuConverter.
    Returns imu = (imuTime, imuRotX, imuRotY, imuRotZ) for projectPointCloud, imuAvailable, n_popped (messages the reference
    pops from the queue front), and imuRollInit / imuPitchInit / imuYawInit (None when no message precedes the scan start).
    Host-only helper of the library (b2_imu_deskew_info): the table is a serial double recurrence of a few dozen terms."""
    stamp, quat, gyro = imuQueue
    stamp = np.ascontiguousarray(stamp, np.float64).reshape(-1)
    gyro = np.ascontiguousarray(gyro, np.float64).reshape(-1, 3)
    quat = None if quat is None else np.ascontiguousarray(quat, np.float64).reshape(-1, 4)
    n = len(stamp)
    if len(gyro) != n or (quat is not None and len(quat) != n):
        raise ValueError("imuQueue arrays differ in length")
    t, rx, ry, rz = (np.zeros(capacity, np.float64) for _ in range(4))
    nt, npop, avail = C.c_int(0), C.c_int(0), C.c_int(0)
    rpy = np.full(3, np.nan, np.float32)
    capi.check(capi.lib().b2_imu_deskew_info(capi.ptr(stamp), capi.ptr(quat), capi.ptr(gyro), n, float(timeScanCur), float(timeScanEnd),
                                             capi.ptr(t), capi.ptr(rx), capi.ptr(ry), capi.ptr(rz), capacity,
                                             C.byref(nt), C.byref(npop), C.byref(avail), capi.ptr(rpy)))
    k = nt.value
    has_rpy = not np.isnan(rpy[0])
    return dict(imu=(t[:k].copy(), rx[:k].copy(), ry[:k].copy(), rz[:k].copy()), imuAvailable=bool(avail.value), n_popped=npop.value,
                imuRollInit=float(rpy[0]) if has_rpy else None, imuPitchInit=float(rpy[1]) if has_rpy else None,
                imuYawInit=float(rpy[2]) if has_rpy else None)


def cloud_info_meta(seq=0, stamp=(0, 0), frame_id="", lidarFrame="", imuAvailable=0, odomAvailable=0, imuRollInit=0.0,
                    imuPitchInit=0.0, imuYawInit=0.0, initialGuess=(0.0,) * 6):
    m = capi.CloudInfoMeta()
    m.seq, m.stamp_sec, m.stamp_nsec = int(seq), int(stamp[0]), int(stamp[1])
    m.frame_id, m.cloud_frame_id = frame_id.encode(), lidarFrame.encode()
    m.imu_available, m.odom_available = int(imuAvailable), int(odomAvailable)
    m.imu_roll_init, m.imu_pitch_init, m.imu_yaw_init = imuRollInit, imuPitchInit, imuYawInit
    (m.initial_guess_x, m.initial_guess_y, m.initial_guess_z, m.initial_guess_roll, m.initial_guess_pitch,
     m.initial_guess_yaw) = initialGuess
    return m


def parse_cloud_info(msg):
    """Host-only: b2_cloud_info_parse on a serialised lio_sam/cloud_info. Returns a dict with the message's field names;
    clouds come back as (n, point_step / 4) float32 arrays of the raw point records plus their field offsets."""
    raw = np.ascontiguousarray(np.frombuffer(msg, np.uint8))
    v = capi.CloudInfoView()
    capi.check(capi.lib().b2_cloud_info_parse(capi.ptr(raw), raw.size, C.byref(v)))
    base = raw.ctypes.data

    def arr(p, n, dt):
        return np.zeros(0, dt) if not n else raw[p - base:p - base + 4 * n].copy().view(dt)

    def cloud(c):
        n = c.width * c.height
        pts = (raw[c.data - base:c.data - base + n * c.point_step].copy().view(np.float32).reshape(n, c.point_step // 4)
               if n else np.zeros((0, max(c.point_step // 4, 1)), np.float32))
        return dict(points=pts, width=c.width, height=c.height, point_step=c.point_step, row_step=c.row_step, n_fields=c.n_fields,
                    offsets=(c.off_x, c.off_y, c.off_z, c.off_intensity), is_dense=c.is_dense, is_bigendian=c.is_bigendian)

    return dict(seq=v.seq, stamp=(v.stamp_sec, v.stamp_nsec),
                frame_id=bytes(raw[v.frame_id - base:v.frame_id - base + v.frame_id_len]).decode() if v.frame_id_len else "",
                startRingIndex=arr(v.start_ring_index, v.n_start_ring_index, np.int32),
                endRingIndex=arr(v.end_ring_index, v.n_end_ring_index, np.int32),
                pointColInd=arr(v.point_col_ind, v.n_point_col_ind, np.int32),
                pointRange=arr(v.point_range, v.n_point_range, np.float32),
                imuAvailable=v.imu_available, odomAvailable=v.odom_available,
                imuRollInit=v.imu_roll_init, imuPitchInit=v.imu_pitch_init, imuYawInit=v.imu_yaw_init,
                initialGuess=(v.initial_guess_x, v.initial_guess_y, v.initial_guess_z, v.initial_guess_roll,
                              v.initial_guess_pitch, v.initial_guess_yaw),
                **{k: cloud(getattr(v, k)) for k in ("cloud_deskewed", "cloud_corner", "cloud_surface", "key_frame_cloud",
                                                      "key_frame_color", "key_frame_poses", "key_frame_map")})
