"""Host-side mirror of the reference's interfaces for the scan-registration path, on top of the C ABI.

Class and method names follow the seams listed in SURVEY.md §8b so call sites read like the reference:
  VoxelGrid           <- pcl::VoxelGrid<PointT>            (featureExtraction.cpp:233-234, mapOptmization.cpp:955-967, ...)
  KdTreeFLANN         <- pcl::KdTreeFLANN<PointType>       (mapOptmization.cpp:1289-1290, :987, :1079), batched
  ScanToMapOptimizer  <- mapOptimization::{cornerOptimization, surfOptimization, combineOptimizationCoeffs,
                         LMOptimization, scan2MapOptimization}  (mapOptmization.cpp:974-1310)
Clouds are numpy float32 arrays, (n, 4) packed x,y,z,intensity or (n, 8) in PCL's 32-byte PointXYZI layout.
Everything computes on the GPU through libb2reg.so; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import capi


class VoxelGrid:
    def __init__(self):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2_voxel_create(C.byref(self._h)))
        self._cloud = None

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_voxel_destroy(self._h)
        except Exception:
            pass

    def setLeafSize(self, lx, ly, lz):
        capi.check(capi.lib().b2_voxel_set_leaf_size(self._h, lx, ly, lz))

    def setMinimumPointsNumberPerVoxel(self, n):
        capi.check(capi.lib().b2_voxel_set_min_points_per_voxel(self._h, int(n)))

    def setInputCloud(self, cloud):
        self._cloud = cloud

    def filter(self, return_voxel_index=False):
        """Returns the downsampled cloud (same column layout as the input). PCL semantics: empty in -> empty out;
        index overflow -> the input is returned unchanged (self.refused = True)."""
        if self._cloud is None:
            raise RuntimeError("VoxelGrid.filter: no input cloud")
        pts, stride = capi.as_points(self._cloud)
        n, cols = pts.shape
        n_fields = 3 if cols == 3 else 4
        out = np.zeros((max(n, 1), cols), np.float32)
        n_out = C.c_size_t(0)
        refused = C.c_int(0)
        vop = np.empty(max(n, 1), np.int32) if return_voxel_index else None
        capi.check(capi.lib().b2_voxel_filter(self._h, capi.ptr(pts), stride, n, n_fields, capi.ptr(out), cols * 4,
                                              out.shape[0], C.byref(n_out), C.byref(refused), capi.ptr(vop)))
        self.refused = bool(refused.value)
        res = out[:n_out.value].copy()
        return (res, vop[:n]) if return_voxel_index else res


def _voxel_last_gpu_ms(self):
    ms = C.c_float(0)
    capi.check(capi.lib().b2_voxel_last_gpu_ms(self._h, C.byref(ms)))
    return ms.value


VoxelGrid.lastGpuMs = _voxel_last_gpu_ms


class KdTreeFLANN:
    """Batched drop-in for the kd-tree queries of the LM loop. Exact for neighbours closer than max_dist."""

    def __init__(self, max_dist=1.0):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2_knn_create(C.byref(self._h), max_dist))

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_knn_destroy(self._h)
        except Exception:
            pass

    def setInputCloud(self, cloud):
        pts, stride = capi.as_points(cloud)
        self._keep = pts
        capi.check(capi.lib().b2_knn_set_input_cloud(self._h, capi.ptr(pts), stride, pts.shape[0]))

    def nearestKSearch(self, queries, k):
        """queries (m, >=3) -> (indices (m,k) int32, sq_dists (m,k) float32), ascending (distance, index)."""
        q, stride = capi.as_points(queries)
        m = q.shape[0]
        idx = np.empty((m, k), np.int32)
        d2 = np.empty((m, k), np.float32)
        capi.check(capi.lib().b2_knn_nearest_k_search(self._h, capi.ptr(q), stride, m, k, capi.ptr(idx), capi.ptr(d2)))
        return idx, d2


class ScanToMapOptimizer:
    """Members named after the reference's: transformTobeMapped, isDegenerate, matP, laserCloud*DS."""

    def __init__(self, max_batch=1, **overrides):
        L = capi.lib()
        self.params = capi.S2MParams()
        L.b2_s2m_default_params(C.byref(self.params))
        self.params.max_batch = max_batch
        for k, v in overrides.items():
            setattr(self.params, k, v)
        self._h = C.c_void_p()
        capi.check(L.b2_s2m_create(C.byref(self._h), C.byref(self.params)))
        self.transformTobeMapped = np.zeros(6, np.float32)
        self.isDegenerate = False
        self.matP = np.zeros((6, 6), np.float32)
        self._nc = self._ns = 0

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_s2m_destroy(self._h)
        except Exception:
            pass

    # kdtreeCornerFromMap->setInputCloud(laserCloudCornerFromMapDS); kdtreeSurfFromMap->setInputCloud(...)
    def setInputMap(self, laserCloudCornerFromMapDS, laserCloudSurfFromMapDS):
        c, cs = capi.as_points(laserCloudCornerFromMapDS)
        s, ss = capi.as_points(laserCloudSurfFromMapDS)
        capi.check(capi.lib().b2_s2m_set_map(self._h, capi.ptr(c), cs, c.shape[0], capi.ptr(s), ss, s.shape[0]))

    def setInputMapFromLocalMap(self, local_map):
        """The same two setInputCloud calls on the device-resident result of LocalMap.extractCloud (no host copy)."""
        capi.check(capi.lib().b2_s2m_set_map_from_localmap(self._h, local_map._h))

    def rebuildMapIndex(self):
        """kdtree*FromMap->setInputCloud again on the device-resident map clouds (mapOptmization.cpp:1289-1290 runs per scan)."""
        capi.check(capi.lib().b2_s2m_rebuild_map_index(self._h))

    def lastStepGpuMs(self):
        """Device ms from the start of the last rebuildMapIndex to the end of the solve that followed it."""
        ms = C.c_float(0)
        capi.check(capi.lib().b2_s2m_last_step_gpu_ms(self._h, C.byref(ms)))
        return ms.value

    def setInputScan(self, laserCloudCornerLastDS, laserCloudSurfLastDS):
        c, cs = capi.as_points(laserCloudCornerLastDS, 4)
        s, ss = capi.as_points(laserCloudSurfLastDS, 4)
        self._nc, self._ns = c.shape[0], s.shape[0]
        capi.check(capi.lib().b2_s2m_set_scan(self._h, capi.ptr(c), cs, c.shape[0], capi.ptr(s), ss, s.shape[0]))

    def _scan_counts(self, nc, ns):
        self._nc, self._ns = nc.value, ns.value
        self.laserCloudCornerLastDSNum, self.laserCloudSurfLastDSNum = nc.value, ns.value

    def setInputScanDownsampled(self, laserCloudCornerLast, laserCloudSurfLast, downSizeFilterCorner, downSizeFilterSurf):
        """downsampleCurrentScan (mapOptmization.cpp:940-958): both VoxelGrids on the device, the result never leaves it."""
        c, cs = capi.as_points(laserCloudCornerLast, 4)
        s, ss = capi.as_points(laserCloudSurfLast, 4)
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib().b2_s2m_set_scan_downsampled(self._h, downSizeFilterCorner._h, capi.ptr(c), cs, c.shape[0],
                                                          downSizeFilterSurf._h, capi.ptr(s), ss, s.shape[0], C.byref(nc), C.byref(ns)))
        self._scan_counts(nc, ns)

    def laserCloudInfoHandler(self, msg, downSizeFilterCorner, downSizeFilterSurf):
        """mapOptimization::laserCloudInfoHandler (:237-247) + downsampleCurrentScan on a serialised feature/cloud_info."""
        from .frontend import parse_cloud_info
        info = parse_cloud_info(msg)
        self.setInputScanDownsampled(info["cloud_corner"]["points"], info["cloud_surface"]["points"], downSizeFilterCorner, downSizeFilterSurf)
        return info

    def setInputScanFromFrontEnd(self, front_end, downSizeFilterCorner, downSizeFilterSurf):
        """The same hand-off inside one process: cornerCloud / surfaceCloud of ScanFrontEnd.extractFeatures stay in HBM."""
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib().b2_s2m_set_scan_from_front_end(self._h, front_end._h, downSizeFilterCorner._h, downSizeFilterSurf._h,
                                                             C.byref(nc), C.byref(ns)))
        self._scan_counts(nc, ns)

    def getInputScan(self):
        """laserCloudCornerLastDS, laserCloudSurfLastDS as the optimiser holds them."""
        out = []
        for which in (0, 1):
            n = C.c_size_t(0)
            capi.check(capi.lib().b2_s2m_get_scan(self._h, which, None, 0, C.byref(n)))
            a = np.empty((n.value, 4), np.float32)
            capi.check(capi.lib().b2_s2m_get_scan(self._h, which, capi.ptr(a), n.value, C.byref(n)))
            out.append(a)
        return out

    def setInputScanBatch(self, corners, surfs):
        """corners / surfs: lists of (n_i, 4) clouds, one per scan."""
        co = np.zeros(len(corners) + 1, np.int32)
        so = np.zeros(len(surfs) + 1, np.int32)
        co[1:] = np.cumsum([len(c) for c in corners])
        so[1:] = np.cumsum([len(s) for s in surfs])
        c = np.ascontiguousarray(np.concatenate(corners), np.float32) if co[-1] else np.zeros((0, 4), np.float32)
        s = np.ascontiguousarray(np.concatenate(surfs), np.float32) if so[-1] else np.zeros((0, 4), np.float32)
        self._batch = len(corners)
        capi.check(capi.lib().b2_s2m_set_scan_batch(self._h, len(corners), capi.ptr(c), 16, capi.ptr(co), capi.ptr(s), 16, capi.ptr(so)))

    def setState(self, isDegenerate, matP=None):
        m = None if matP is None else np.ascontiguousarray(matP, np.float32).reshape(36)
        capi.check(capi.lib().b2_s2m_set_state(self._h, int(isDegenerate), capi.ptr(m)))

    def LMIteration(self, iterCount):
        """cornerOptimization + surfOptimization + combineOptimizationCoeffs + LMOptimization(iterCount).
        Returns the bool LMOptimization returns; updates transformTobeMapped / isDegenerate / matP in place."""
        pose = np.ascontiguousarray(self.transformTobeMapped, np.float32)
        n_sel, ran, conv, deg = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        matP = np.zeros(36, np.float32)
        capi.check(capi.lib().b2_s2m_iterate(self._h, capi.ptr(pose), iterCount, C.byref(n_sel), C.byref(ran), C.byref(conv),
                                             C.byref(deg), capi.ptr(matP)))
        self.transformTobeMapped = pose
        self.laserCloudSelNum = n_sel.value
        self.ran = bool(ran.value)
        self.isDegenerate = bool(deg.value)
        self.matP = matP.reshape(6, 6)
        return bool(conv.value)

    def scan2MapOptimization(self, max_iterations=30, record_history=False, want_matP=True):
        """The whole loop on the device. Returns dict(iters, converged, not_enough, pose_history).
        want_matP=False lets iteration 0 skip the 6x6 eigen-decomposition when the normal matrix is certified
        non-degenerate (matP is only ever read when isDegenerate, mapOptmization.cpp:1253-1258)."""
        pose = np.ascontiguousarray(self.transformTobeMapped, np.float32)
        it, conv, deg, ne = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        matP = np.zeros(36, np.float32) if want_matP else None
        hist = np.zeros((max_iterations, 6), np.float32) if record_history else None
        capi.check(capi.lib().b2_s2m_solve(self._h, capi.ptr(pose), max_iterations, C.byref(it), C.byref(conv), C.byref(deg),
                                           capi.ptr(matP), C.byref(ne), capi.ptr(hist)))
        self.transformTobeMapped = pose
        if not ne.value:
            self.isDegenerate = bool(deg.value)
            if matP is not None:
                self.matP = matP.reshape(6, 6)
        return dict(iters=it.value, converged=bool(conv.value), not_enough=bool(ne.value),
                    pose_history=None if hist is None else hist[:it.value])

    def scan2MapOptimizationBatch(self, poses, max_iterations=30):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6).copy()
        B = poses.shape[0]
        it = np.zeros(B, np.int32)
        conv = np.zeros(B, np.int32)
        deg = np.zeros(B, np.int32)
        capi.check(capi.lib().b2_s2m_solve_batch(self._h, capi.ptr(poses), max_iterations, capi.ptr(it), capi.ptr(conv), capi.ptr(deg)))
        return dict(poses=poses, iters=it, converged=conv.astype(bool), degenerate=deg.astype(bool))

    def lastGpuMs(self):
        ms, n = C.c_float(0), C.c_int(0)
        capi.check(capi.lib().b2_s2m_last_gpu_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def countCandidates(self, enable):
        """Measurement hook: batched solves count the map points their search loads; returns the last counted total."""
        n = C.c_ulonglong(0)
        capi.check(capi.lib().b2_s2m_count_candidates(self._h, 1 if enable else 0, C.byref(n)))
        return n.value

    def getPass(self, which):
        n = self._ns if which else self._nc
        idx = np.empty((n, 5), np.int32)
        d2 = np.empty((n, 5), np.float32)
        coeff = np.empty((n, 4), np.float32)
        flag = np.empty(n, np.uint8)
        capi.check(capi.lib().b2_s2m_get_pass(self._h, which, capi.ptr(idx), capi.ptr(d2), capi.ptr(coeff), capi.ptr(flag)))
        return dict(idx=idx, d2=d2, coeff=coeff, flag=flag)

    def getNormalEquations(self):
        AtA, AtB, X = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32)
        capi.check(capi.lib().b2_s2m_get_normal_equations(self._h, capi.ptr(AtA), capi.ptr(AtB), capi.ptr(X)))
        return AtA.reshape(6, 6), AtB, X


class LocalMap:
    """Key-frame store + extractCloud of mapOptimization (mapOptmization.cpp:899-938), device-resident.
    Members named after the reference's: cornerCloudKeyFrames / surfCloudKeyFrames / cloudKeyPoses6D live in the handle,
    laserCloud{Corner,Surf}FromMap[DS] are read back with get()."""

    def __init__(self, mappingCornerLeafSize=0.2, mappingSurfLeafSize=0.4, surroundingKeyframeSearchRadius=50.0):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2_localmap_create(C.byref(self._h), float(mappingCornerLeafSize), float(mappingSurfLeafSize)))
        self.surroundingKeyframeSearchRadius = float(surroundingKeyframeSearchRadius)
        self.cloudKeyPoses6D = []

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_localmap_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def saveKeyFrame(self, thisCornerKeyFrame, thisSurfKeyFrame, pose6):
        """cornerCloudKeyFrames.push_back / surfCloudKeyFrames.push_back / cloudKeyPoses6D.push_back (:1512-1524)."""
        c, cs = capi.as_points(thisCornerKeyFrame, 4)
        s, ss = capi.as_points(thisSurfKeyFrame, 4)
        p = np.ascontiguousarray(pose6, np.float32)
        k = C.c_int()
        capi.check(capi.lib().b2_localmap_add_keyframe(self._h, capi.ptr(c), cs, len(c), capi.ptr(s), ss, len(s), capi.ptr(p), C.byref(k)))
        self.cloudKeyPoses6D.append(p.copy())
        return k.value

    def correctPose(self, key_index, pose6):
        p = np.ascontiguousarray(pose6, np.float32)
        capi.check(capi.lib().b2_localmap_set_pose(self._h, int(key_index), capi.ptr(p)))
        self.cloudKeyPoses6D[key_index] = p.copy()

    def clearMapContainer(self):
        """laserCloudMapContainer.clear() (:1591)."""
        capi.check(capi.lib().b2_localmap_clear_cache(self._h))

    def extractCloud(self, cloudToExtract):
        """cloudToExtract: key indices in visiting order. Applies the distance test of :905 against the latest key pose,
        then assembles and downsamples on the device. Returns (n_corner_ds, n_surf_ds)."""
        last = self.cloudKeyPoses6D[-1][3:6]
        keep = [int(k) for k in cloudToExtract
                if np.sqrt(np.float32(np.sum((self.cloudKeyPoses6D[int(k)][3:6] - last) ** 2, dtype=np.float32))) <= self.surroundingKeyframeSearchRadius]
        idx = np.ascontiguousarray(keep, np.int32)
        nc, ns = C.c_size_t(), C.c_size_t()
        capi.check(capi.lib().b2_localmap_extract(self._h, capi.ptr(idx), len(idx), C.byref(nc), C.byref(ns)))
        return nc.value, ns.value

    def get(self, which):
        """which: 'corner', 'surf' (laserCloud*FromMap) or 'cornerDS', 'surfDS' (laserCloud*FromMapDS)."""
        w = {"corner": 0, "surf": 1, "cornerDS": 2, "surfDS": 3}[which]
        n = C.c_size_t()
        capi.check(capi.lib().b2_localmap_get(self._h, w, None, 16, 0, C.byref(n)))
        out = np.empty((n.value, 4), np.float32)
        if n.value:
            capi.check(capi.lib().b2_localmap_get(self._h, w, capi.ptr(out), 16, n.value, C.byref(n)))
        return out

    def lastGpuMs(self):
        ms, nc = C.c_float(), C.c_size_t()
        capi.check(capi.lib().b2_localmap_last_gpu_ms(self._h, C.byref(ms), C.byref(nc)))
        return ms.value, nc.value


def transformPointCloud(cloud, transformIn):
    """mapOptimization::transformPointCloud (mapOptmization.cpp:286-305); transformIn = (roll,pitch,yaw,x,y,z)."""
    pts, stride = capi.as_points(cloud, 4)
    out = np.zeros_like(pts)
    pose = np.ascontiguousarray(transformIn, np.float32)
    capi.check(capi.lib().b2_transform_cloud(capi.ptr(pts), stride, pts.shape[0], capi.ptr(pose), capi.ptr(out), pts.shape[1] * 4))
    return out


class ICPRegistrator:
    """The yaw grid search of SensorsCalibration's lidar2lidar auto-calibration (auto_calib/src/registration_icp.cpp:49-100),
    member names as there. Clouds are (n, >=3) float32 (the non-ground clouds tgt_ngcloud_ / src_ngcloud_)."""

    def __init__(self):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2_nnerr_create(C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_nnerr_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def SetTargetCloud(self, ngcloud):
        p, st = capi.as_points(ngcloud, 3)
        capi.check(capi.lib().b2_nnerr_set_target(self._h, capi.ptr(p), st, len(p)))

    def SetSourceCloud(self, ngcloud):
        p, st = capi.as_points(ngcloud, 3)
        capi.check(capi.lib().b2_nnerr_set_source(self._h, capi.ptr(p), st, len(p)))

    def CalculateICPError(self, T):
        T = np.ascontiguousarray(T, np.float64)
        s, n = C.c_double(), C.c_size_t()
        capi.check(capi.lib().b2_nnerr_evaluate(self._h, capi.ptr(T), C.byref(s), C.byref(n)))
        return s.value

    def RegistrationByICP(self, init_guess):
        g = np.ascontiguousarray(init_guess, np.float64)
        T = np.empty((4, 4), np.float64)
        yaw, err, ev = C.c_double(), C.c_double(), C.c_int()
        capi.check(capi.lib().b2_nnerr_yaw_search(self._h, capi.ptr(g), capi.ptr(T), C.byref(yaw), C.byref(err), C.byref(ev)))
        ms = C.c_float()
        capi.check(capi.lib().b2_nnerr_last_gpu_ms(self._h, C.byref(ms)))
        return dict(transform=T, best_yaw=yaw.value, min_error=err.value, evaluations=ev.value, gpu_ms=ms.value)


class IterativeClosestPoint:
    """pcl::IterativeClosestPoint<PointType, PointType> as the loop-closure thread uses it (mapOptmization.cpp:559-586);
    method names as PCL's. Clouds are (n, >=3) float32."""

    def __init__(self):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2_icp_create(C.byref(self._h)))
        self._n_src = 0

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_icp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def setMaxCorrespondenceDistance(self, d):
        capi.check(capi.lib().b2_icp_set_max_correspondence_distance(self._h, float(d)))

    def setMaximumIterations(self, n):
        capi.check(capi.lib().b2_icp_set_maximum_iterations(self._h, int(n)))

    def setTransformationEpsilon(self, e):
        capi.check(capi.lib().b2_icp_set_transformation_epsilon(self._h, float(e)))

    def setEuclideanFitnessEpsilon(self, e):
        capi.check(capi.lib().b2_icp_set_euclidean_fitness_epsilon(self._h, float(e)))

    def setRANSACIterations(self, n):
        capi.check(capi.lib().b2_icp_set_ransac_iterations(self._h, int(n)))

    def setInputSource(self, cloud):
        p, st = capi.as_points(cloud, 3)
        capi.check(capi.lib().b2_icp_set_input_source(self._h, capi.ptr(p), st, len(p)))
        self._n_src = len(p)

    def setInputTarget(self, cloud):
        p, st = capi.as_points(cloud, 3)
        capi.check(capi.lib().b2_icp_set_input_target(self._h, capi.ptr(p), st, len(p)))

    def align(self, guess=None, want_output=False):
        g = None if guess is None else np.ascontiguousarray(guess, np.float32)
        out = np.zeros((self._n_src, 3), np.float32) if want_output else None
        capi.check(capi.lib().b2_icp_align(self._h, capi.ptr(g), capi.ptr(out), 12))
        return out

    def hasConverged(self):
        v = C.c_int()
        capi.check(capi.lib().b2_icp_has_converged(self._h, C.byref(v)))
        return bool(v.value)

    def getFitnessScore(self):
        v = C.c_double()
        capi.check(capi.lib().b2_icp_get_fitness_score(self._h, C.byref(v)))
        return v.value

    def getFinalTransformation(self):
        T = np.empty((4, 4), np.float32)
        capi.check(capi.lib().b2_icp_get_final_transformation(self._h, capi.ptr(T)))
        return T

    def getFinalNumIteration(self):
        v = C.c_int()
        capi.check(capi.lib().b2_icp_get_final_num_iteration(self._h, C.byref(v)))
        return v.value

    def lastGpuMs(self):
        ms, n = C.c_float(), C.c_int()
        capi.check(capi.lib().b2_icp_last_gpu_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class FusionPc:
    """The fusion callback of PointClouds_Fusion (fusion_pointclouds.cpp:55-115): transform the child clouds into the parent
    frame, concatenate in the reference's order, external pass-through box, internal conditional-removal box."""

    def __init__(self):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2_fusion_create(C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                capi.lib().b2_fusion_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def fuse(self, clouds, transforms, external_bounds=None, internal_bounds=None):
        """clouds: list of (n, >=4) float32 in fusion order; transforms: list of 4x4 (or None) of the same length;
        bounds: ((xmin, ymin, zmin), (xmax, ymax, zmax)) or None. Returns the fused (m, 4) xyzi cloud."""
        L = capi.lib()
        capi.check(L.b2_fusion_clear(self._h))
        for c, T in zip(clouds, transforms):
            p, st = capi.as_points(c, 4)
            t = None if T is None else np.ascontiguousarray(T, np.float64)
            capi.check(L.b2_fusion_add_cloud(self._h, capi.ptr(p), st, len(p), capi.ptr(t)))
        for fn, b in ((L.b2_fusion_set_external_bounds, external_bounds), (L.b2_fusion_set_internal_bounds, internal_bounds)):
            if b is None:
                capi.check(fn(self._h, 0, None, None))
            else:
                lo, hi = np.ascontiguousarray(b[0], np.float64), np.ascontiguousarray(b[1], np.float64)
                capi.check(fn(self._h, 1, capi.ptr(lo), capi.ptr(hi)))
        nf, no = C.c_size_t(), C.c_size_t()
        capi.check(L.b2_fusion_run(self._h, C.byref(nf), C.byref(no)))
        out = np.empty((no.value, 4), np.float32)
        if no.value:
            capi.check(L.b2_fusion_get(self._h, capi.ptr(out), 16, no.value, C.byref(no)))
        self.n_fused = nf.value
        return out

    def lastGpuMs(self):
        ms = C.c_float()
        capi.check(capi.lib().b2_fusion_last_gpu_ms(self._h, C.byref(ms)))
        return ms.value
