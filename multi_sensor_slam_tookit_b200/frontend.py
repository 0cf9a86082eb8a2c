"""Host-side mirror of LIO-SAM's per-scan front end on top of the C ABI (b2_scan_*).

  ScanFrontEnd.projectPointCloud  <- ImageProjection::projectPointCloud + deskewPoint + cloudExtraction
                                     (liosam_ws/src/LIO-SAM/src/imageProjection.cpp:446-598)
  ScanFrontEnd.extractFeatures    <- FeatureExtraction::calculateSmoothness + markOccludedPoints + extractFeatures
                                     (liosam_ws/src/LIO-SAM/src/featureExtraction.cpp:81-238)
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
