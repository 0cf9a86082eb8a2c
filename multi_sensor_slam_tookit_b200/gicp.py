"""Host mirror of the Open3D interface Multi_LiCa uses for its GICP calibration, on top of the C ABI (libb2reg.so).

Reference call sites (under /root/reference/Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/):
  calibration/Calibration.py:306-307  o3d.geometry.PointCloud(self.source.pcd)
  calibration/Calibration.py:314-315  pcd.voxel_down_sample(voxel_size)
  calibration/Calibration.py:327-328  pcd.estimate_normals()
  calibration/Calibration.py:331-340  o3d.pipelines.registration.registration_generalized_icp(...)
  multi_lidar_calibrator.py:202-219, 302-321   one independent Calibration per (source, target) pair
The names and argument meaning follow Open3D so the parity tests read like Calibration.py. Clouds live in HBM; numpy
arrays cross the boundary only in `PointCloud(points)` / `.points` / `.normals`. There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import capi


class PointCloud:
    """o3d.geometry.PointCloud restricted to what Calibration.py touches: points, normals, copy-construction."""

    def __init__(self, points=None, _handle=None):
        L = capi.lib()
        if _handle is not None:
            self._h = _handle
            return
        h = C.c_void_p()
        capi.check(L.b2_cloud_create(C.byref(h)))
        self._h = h
        if isinstance(points, PointCloud):            # o3d.geometry.PointCloud(other) copies
            self._set_points(points.points)
            if points.has_normals():
                self.normals = points.normals
        elif points is not None:
            self._set_points(points)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                capi.lib().b2_cloud_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _set_points(self, pts):
        pts = np.asarray(pts)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise ValueError(f"expected (n, 3) points, got {pts.shape}")
        L = capi.lib()
        if pts.dtype == np.float32:
            p = np.ascontiguousarray(pts)
            capi.check(L.b2_cloud_set_points_f32(self._h, capi.ptr(p), 12, len(p)))
        else:
            p = np.ascontiguousarray(pts, dtype=np.float64)
            capi.check(L.b2_cloud_set_points(self._h, capi.ptr(p), len(p)))

    def __len__(self):
        n = C.c_size_t()
        capi.check(capi.lib().b2_cloud_size(self._h, C.byref(n), None))
        return n.value

    def has_normals(self):
        hn = C.c_int()
        capi.check(capi.lib().b2_cloud_size(self._h, None, C.byref(hn)))
        return bool(hn.value)

    @property
    def points(self):
        out = np.empty((len(self), 3), np.float64)
        capi.check(capi.lib().b2_cloud_get_points(self._h, capi.ptr(out)))
        return out

    @points.setter
    def points(self, pts):
        self._set_points(pts)

    @property
    def normals(self):
        out = np.empty((len(self), 3), np.float64)
        capi.check(capi.lib().b2_cloud_get_normals(self._h, capi.ptr(out)))
        return out

    @normals.setter
    def normals(self, nrm):
        nrm = np.ascontiguousarray(nrm, dtype=np.float64)
        if nrm.shape != (len(self), 3):
            raise ValueError("normals must be (n, 3)")
        capi.check(capi.lib().b2_cloud_set_normals(self._h, capi.ptr(nrm)))

    def voxel_down_sample(self, voxel_size, return_voxel_rank=False):
        """Calibration.py:314-315. Raises like Open3D on voxel_size <= 0."""
        if not voxel_size > 0.0:
            raise RuntimeError("[Open3D Error] voxel_size <= 0.")
        out = C.c_void_p()
        rank = np.empty(len(self), np.int32) if return_voxel_rank else None
        capi.check(capi.lib().b2_cloud_voxel_down_sample(self._h, float(voxel_size), C.byref(out), capi.ptr(rank)))
        pc = PointCloud(_handle=out)
        return (pc, rank) if return_voxel_rank else pc

    @classmethod
    def from_host_sharded(cls, points, comm):
        """Every rank passes the same (n, 3) float64 array; each uploads 1/world of it and the slices are all-gathered."""
        pts = np.ascontiguousarray(points, dtype=np.float64)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise ValueError("points must be (n, 3)")
        self = cls()
        capi.check(capi.lib().b2_cloud_set_points_sharded(self._h, capi.ptr(pts), pts.shape[0], comm._h if comm is not None else None))
        return self

    def estimate_normals_sharded(self, comm, knn=30):
        """estimate_normals() with the 30-NN work split over the ranks of comm (identical clouds on every rank)."""
        capi.check(capi.lib().b2_cloud_estimate_normals_sharded(self._h, int(knn), comm._h if comm is not None else None))

    def estimate_normals(self, knn=30):
        """Calibration.py:327-328 (Open3D default search parameter: KDTreeSearchParamKNN(knn=30))."""
        capi.check(capi.lib().b2_cloud_estimate_normals(self._h, int(knn)))
        return True

    def transform(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64)
        if T.shape != (4, 4):
            raise ValueError("transformation must be 4x4")
        capi.check(capi.lib().b2_cloud_transform(self._h, capi.ptr(T)))
        return self

    def lastGpuMs(self):
        ms = C.c_float()
        capi.check(capi.lib().b2_cloud_last_gpu_ms(self._h, C.byref(ms)))
        return ms.value


class TransformationEstimationForGeneralizedICP:
    def __init__(self, epsilon=1e-3):
        self.epsilon = float(epsilon)


class ICPConvergenceCriteria:
    def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
        self.relative_fitness = float(relative_fitness)
        self.relative_rmse = float(relative_rmse)
        self.max_iteration = int(max_iteration)


class RegistrationResult:
    def __init__(self):
        self.transformation = np.eye(4)
        self.fitness = 0.0
        self.inlier_rmse = 0.0
        self.correspondence_set = np.empty((0, 2), np.int32)
        self.iterations = 0
        self.converged = False
        self.gpu_ms = 0.0
        self.gpu_launches = 0
        self.fitness_history = np.empty(0)
        self.rmse_history = np.empty(0)
        self.evaluation_ms = np.empty(0, np.float32)

    def __repr__(self):
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, "
                f"and correspondence_set size of {len(self.correspondence_set)}")


class Communicator:
    """NCCL communicator of the sharded registration (one process per GPU). `id_bytes`: the 128 bytes rank 0 got from
    Communicator.unique_id(), shipped by the host program (torch.distributed broadcast in bench.py)."""

    def __init__(self, id_bytes, rank, world):
        buf = np.frombuffer(bytes(id_bytes), dtype=np.uint8).copy()
        if buf.size != 128:
            raise ValueError("NCCL unique id is 128 bytes")
        h = C.c_void_p()
        capi.check(capi.lib().b2_comm_create(C.byref(h), capi.ptr(buf), int(rank), int(world)))
        self._h, self.rank, self.world = h, int(rank), int(world)

    @staticmethod
    def unique_id():
        buf = np.zeros(128, np.uint8)
        capi.check(capi.lib().b2_comm_unique_id(capi.ptr(buf)))
        return buf.tobytes()

    def allreduce(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        capi.check(capi.lib().b2_comm_allreduce_f64(self._h, capi.ptr(v), v.size))
        return v

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                capi.lib().b2_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass


SHARD_BLOCK_POINTS = 4096


def shard_blocks(n, rank, world, block=SHARD_BLOCK_POINTS):
    """[begin, end) ranges of the n (cell-sorted) source points `rank` evaluates: blocks of `block` consecutive points dealt
    round-robin, block b to rank b % world (the deal libb2reg uses for the sharded registration)."""
    return [(b * block, min((b + 1) * block, n)) for b in range((n + block - 1) // block) if b % world == rank]


def shard_size(n, rank, world, block=SHARD_BLOCK_POINTS):
    return sum(e - b for b, e in shard_blocks(n, rank, world, block))


def row_slice(n, rank, world):
    """Rows [begin, end) of an n-point cloud that rank owns when the source is sliced by rows."""
    per = (n + world - 1) // world
    b = min(n, per * rank)
    return b, min(n, b + per)


def pair_owner(pair_index, world):
    """Round-robin owner of an independent calibration pair (multi_lidar_calibrator.py:202-219)."""
    return pair_index % world


class GeneralizedICP:
    """Reusable registration object: one target index, many sources / initial guesses."""

    def __init__(self, max_correspondence_distance=1.0, epsilon=0.005, relative_fitness=1e-7, relative_rmse=1e-7,
                 max_iteration=100):
        self._prm = capi.GicpParams(float(max_correspondence_distance), float(epsilon), float(relative_fitness),
                                    float(relative_rmse), int(max_iteration))
        h = C.c_void_p()
        capi.check(capi.lib().b2_gicp_create(C.byref(h), C.byref(self._prm)))
        self._h = h
        self._n_src = 0
        self._comm = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                capi.lib().b2_gicp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def setParams(self, **kw):
        for k, v in kw.items():
            if not hasattr(self._prm, k):
                raise AttributeError(k)
            setattr(self._prm, k, v)
        capi.check(capi.lib().b2_gicp_set_params(self._h, C.byref(self._prm)))

    def setInputTarget(self, cloud):
        capi.check(capi.lib().b2_gicp_set_target(self._h, cloud._h))

    def setInputSource(self, cloud):
        capi.check(capi.lib().b2_gicp_set_source(self._h, cloud._h))
        self._n_src = len(cloud)

    def setInputSourceSlice(self, cloud, begin, end):
        """Index rows [begin, end) of the source only (this rank's slice of a sharded registration); see row_slice()."""
        capi.check(capi.lib().b2_gicp_set_source_slice(self._h, cloud._h, int(begin), int(end)))
        self._n_src = len(cloud)

    def setInputSourceBlocks(self, cloud, rank, world):
        """This rank's blocks of the source's Morton order (4096 points each, dealt round-robin): dense, balanced shards."""
        capi.check(capi.lib().b2_gicp_set_source_blocks(self._h, cloud._h, int(rank), int(world)))
        self._n_src = len(cloud)

    def setShard(self, comm):
        """Shard the source over the ranks of `comm` (None = single GPU)."""
        self._comm = comm
        if comm is None:
            capi.check(capi.lib().b2_gicp_set_shard(self._h, 0, 1, None))
        else:
            capi.check(capi.lib().b2_gicp_set_shard(self._h, comm.rank, comm.world, comm._h))

    def peerHandle(self):
        """CUDA IPC handle (64 bytes) of this rank's exchange area; gather them over the ranks and pass them to setPeers."""
        h = np.zeros(64, np.uint8)
        capi.check(capi.lib().b2_gicp_peer_handle(self._h, capi.ptr(h)))
        return h.tobytes()

    def setPeers(self, rank, world, handles):
        """handles: the peerHandle() bytes of all ranks in rank order. Returns False (and keeps the NCCL all-reduce) when the
        GPUs cannot map each other's memory."""
        buf = np.frombuffer(b"".join(handles), np.uint8).copy()
        try:
            capi.check(capi.lib().b2_gicp_set_peers(self._h, int(rank), int(world), capi.ptr(buf)))
            return True
        except capi.B2Error:
            return False

    def setupExchange(self):
        """peerHandle + setPeers in one call: the handles travel over the communicator given to setShard. Returns False (and
        keeps the NCCL all-reduce) when the GPUs cannot map each other's memory."""
        try:
            capi.check(capi.lib().b2_gicp_exchange_setup(self._h))
            return True
        except capi.B2Error:
            return False

    def linearize(self, T, want_correspondences=False):
        T = np.ascontiguousarray(T, dtype=np.float64)
        sums = np.zeros(30, np.float64)
        corr = np.empty(self._n_src, np.int32) if want_correspondences else None
        capi.check(capi.lib().b2_gicp_linearize(self._h, capi.ptr(T), capi.ptr(sums), capi.ptr(corr)))
        return (sums, corr) if want_correspondences else sums

    def align(self, init=None, want_correspondences=True):
        init = np.eye(4) if init is None else np.ascontiguousarray(init, dtype=np.float64)
        if init.shape != (4, 4):
            raise ValueError("init must be 4x4")
        T = np.empty((4, 4), np.float64)
        fit, rmse, it, conv = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        L = capi.lib()
        capi.check(L.b2_gicp_align(self._h, capi.ptr(init), capi.ptr(T), C.byref(fit), C.byref(rmse), C.byref(it), C.byref(conv)))
        r = RegistrationResult()
        r.transformation, r.fitness, r.inlier_rmse = T, fit.value, rmse.value
        r.iterations, r.converged = it.value, bool(conv.value)
        ms, nl = C.c_float(), C.c_int()
        capi.check(L.b2_gicp_last_gpu_ms(self._h, C.byref(ms), C.byref(nl)))
        r.gpu_ms, r.gpu_launches = ms.value, nl.value
        hf, hr, ne = np.zeros(256), np.zeros(256), C.c_int()
        capi.check(L.b2_gicp_get_history(self._h, capi.ptr(hf), capi.ptr(hr), 256, C.byref(ne)))
        r.fitness_history, r.rmse_history = hf[:min(ne.value, 256)].copy(), hr[:min(ne.value, 256)].copy()
        ems, ne2 = np.zeros(256, np.float32), C.c_int()
        capi.check(L.b2_gicp_get_evaluation_ms(self._h, capi.ptr(ems), 256, C.byref(ne2)))
        r.evaluation_ms = ems[:ne2.value].copy()
        if want_correspondences and self._comm is None:
            _, corr = self.linearize(T, want_correspondences=True)
            src = np.nonzero(corr >= 0)[0].astype(np.int32)
            r.correspondence_set = np.stack([src, corr[src]], axis=1)
        return r

    def indexInfo(self):
        h, ppc, b, e = C.c_double(), C.c_double(), C.c_uint32(), C.c_uint32()
        capi.check(capi.lib().b2_gicp_index_info(self._h, C.byref(h), C.byref(ppc), C.byref(b), C.byref(e)))
        return {"target_cell_edge": h.value, "target_points_per_cell": ppc.value, "shard_points": b.value, "shard_block_points": e.value}


def registration_generalized_icp(source, target, max_correspondence_distance, init=None,
                                 estimation_method=None, criteria=None, comm=None):
    """o3d.pipelines.registration.registration_generalized_icp as Calibration.py:331-340 calls it.
    `comm`: optional Communicator — the source is then sharded over its ranks (every rank passes the same clouds)."""
    est = estimation_method or TransformationEstimationForGeneralizedICP()
    crit = criteria or ICPConvergenceCriteria()
    if not (source.has_normals() and target.has_normals()):
        raise RuntimeError("registration_generalized_icp: both clouds need normals (call estimate_normals() first)")
    g = GeneralizedICP(max_correspondence_distance, est.epsilon, crit.relative_fitness, crit.relative_rmse, crit.max_iteration)
    g.setInputTarget(target)
    g.setInputSource(source)
    if comm is not None:
        g.setShard(comm)
    return g.align(init)
