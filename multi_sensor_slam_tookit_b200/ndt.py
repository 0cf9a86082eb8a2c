"""Host mirror of pcl::NormalDistributionsTransform<PointXYZ, PointXYZ> on top of the C ABI (libb2reg.so).

Reference call site: Calibration_Tookit/multi_lidar/src/multi_lidar_calibration/src/multi_lidar_calibrator.cpp:28-72
(PerformNdtOptimize). Method names and argument meaning follow PCL so the parity tests read like that function.
Clouds are (n, 3) float32 arrays (pcl::PointXYZ without the padding word). There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import capi


class NormalDistributionsTransform:
    def __init__(self):
        h = C.c_void_p()
        capi.check(capi.lib().b2_ndt_create(C.byref(h)))
        self._h = h
        self._n_src = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                capi.lib().b2_ndt_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- multi_lidar_calibrator.cpp:37-41
    def setTransformationEpsilon(self, epsilon):
        capi.check(capi.lib().b2_ndt_set_transformation_epsilon(self._h, float(epsilon)))

    def setStepSize(self, step_size):
        capi.check(capi.lib().b2_ndt_set_step_size(self._h, float(step_size)))

    def setResolution(self, resolution):
        capi.check(capi.lib().b2_ndt_set_resolution(self._h, float(resolution)))

    def setMaximumIterations(self, n):
        capi.check(capi.lib().b2_ndt_set_maximum_iterations(self._h, int(n)))

    # ---- :43-44
    def setInputSource(self, cloud):
        p, stride = capi.as_points(cloud, 3)
        capi.check(capi.lib().b2_ndt_set_input_source(self._h, capi.ptr(p), stride, len(p)))
        self._n_src = len(p)

    def setInputTarget(self, cloud):
        p, stride = capi.as_points(cloud, 3)
        capi.check(capi.lib().b2_ndt_set_input_target(self._h, capi.ptr(p), stride, len(p)))

    # ---- :62
    def align(self, guess=None, want_output=False):
        g = np.eye(4, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, dtype=np.float32)
        if g.shape != (4, 4):
            raise ValueError("guess must be 4x4")
        out = np.zeros((self._n_src, 3), np.float32) if want_output else None
        capi.check(capi.lib().b2_ndt_align(self._h, capi.ptr(g), capi.ptr(out), 12))
        return out

    # ---- :64-65, :69
    def hasConverged(self):
        v = C.c_int()
        capi.check(capi.lib().b2_ndt_has_converged(self._h, C.byref(v)))
        return bool(v.value)

    def getFitnessScore(self):
        v = C.c_double()
        capi.check(capi.lib().b2_ndt_get_fitness_score(self._h, C.byref(v)))
        return v.value

    def getTransformationProbability(self):
        v = C.c_double()
        capi.check(capi.lib().b2_ndt_get_transformation_probability(self._h, C.byref(v)))
        return v.value

    def getFinalTransformation(self):
        T = np.empty((4, 4), np.float32)
        capi.check(capi.lib().b2_ndt_get_final_transformation(self._h, capi.ptr(T)))
        return T

    def getFinalNumIteration(self):
        v = C.c_int()
        capi.check(capi.lib().b2_ndt_get_final_num_iteration(self._h, C.byref(v)))
        return v.value

    # ---- parity hooks
    def getVoxels(self):
        L = capi.lib()
        n = C.c_size_t()
        mn, dv = np.empty(3, np.int32), np.empty(3, np.int32)
        capi.check(L.b2_ndt_get_voxels(self._h, 0, C.byref(n), None, None, None, None, None, capi.ptr(mn), capi.ptr(dv)))
        m = max(n.value, 1)
        idx, npts = np.empty(m, np.int32), np.empty(m, np.int32)
        cen, mean, icov = np.empty((m, 3), np.float32), np.empty((m, 3), np.float64), np.empty((m, 9), np.float64)
        capi.check(L.b2_ndt_get_voxels(self._h, m, C.byref(n), capi.ptr(idx), capi.ptr(npts), capi.ptr(cen), capi.ptr(mean),
                                       capi.ptr(icov), capi.ptr(mn), capi.ptr(dv)))
        k = n.value
        return dict(index=idx[:k], npts=npts[:k], centroid=cen[:k], mean=mean[:k], icov=icov[:k].reshape(-1, 3, 3), min_b=mn, div_b=dv)

    def derivatives(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        g, H = np.zeros(6), np.zeros(36)
        s, pairs = C.c_double(), C.c_longlong()
        capi.check(capi.lib().b2_ndt_derivatives(self._h, capi.ptr(p), C.byref(s), capi.ptr(g), capi.ptr(H), C.byref(pairs)))
        return s.value, g, H.reshape(6, 6), pairs.value

    def lastGpuMs(self):
        ms, nl, ne, pr = C.c_float(), C.c_int(), C.c_int(), C.c_longlong()
        capi.check(capi.lib().b2_ndt_last_gpu_ms(self._h, C.byref(ms), C.byref(nl), C.byref(ne), C.byref(pr)))
        return dict(ms=ms.value, launches=nl.value, evaluations=ne.value, pairs_last=pr.value)
