"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes bindings to oracle/liboracle.so (the CPU restatement of the reference hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _opt(ptr_type):
    """ndpointer that also accepts None."""
    base = ptr_type

    class _P(base):
        @classmethod
        def from_param(cls, obj):
            if obj is None:
                return None
            return base.from_param(obj)
    return _P


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    of32, oi32, ou8 = _opt(f32p), _opt(i32p), _opt(u8p)
    L.o_voxel_grid.restype = C.c_int
    L.o_voxel_grid.argtypes = [f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint, f32p, oi32,
                               C.POINTER(C.c_int), oi32]
    L.o_s2m_create.restype = C.c_void_p
    L.o_s2m_create.argtypes = [C.c_int]
    L.o_s2m_destroy.argtypes = [C.c_void_p]
    L.o_s2m_set_map.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
    L.o_s2m_set_scan.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
    L.o_s2m_set_state.argtypes = [C.c_void_p, C.c_int, of32]
    L.o_s2m_get_state.argtypes = [C.c_void_p, C.POINTER(C.c_int), of32]
    L.o_s2m_iterate.restype = C.c_int
    L.o_s2m_iterate.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), of32, of32, of32]
    L.o_s2m_get_pass.argtypes = [C.c_void_p, C.c_int, oi32, of32, of32, ou8]
    L.o_s2m_solve.restype = C.c_int
    L.o_s2m_solve.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                              of32, oi32]
    for name in ("o_knn", "o_knn_brute"):
        fn = getattr(L, name)
        fn.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, i32p, f32p, C.c_int]
    L.o_transform_cloud.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int]
    L.o_eigen3.argtypes = [f32p, f32p, f32p]
    L.o_eigen6.argtypes = [f32p, f32p, f32p]
    L.o_qr_solve6.restype = C.c_int
    L.o_qr_solve6.argtypes = [f32p, f32p, f32p]
    L.o_lu_invert6.restype = C.c_int
    L.o_lu_invert6.argtypes = [f32p, f32p]
    L.o_plane5.argtypes = [f32p, f32p]
    L.o_pose_affine.argtypes = [f32p, f32p]
    L.o_project.restype = C.c_int
    L.o_project.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                            f64p, f64p, f64p, f64p, C.c_int, C.c_double, C.c_int,
                            f32p, f32p, i32p, f32p, i32p, f32p, i32p, i32p]
    L.o_curvature_masks.argtypes = [f32p, i32p, C.c_int, f32p, i32p, i32p]
    L.o_imu_deskew_info.restype = C.c_int
    L.o_imu_deskew_info.argtypes = [f64p, f64p, f64p, C.c_int, C.c_double, C.c_double, f64p, f64p, f64p, f64p, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), f32p]
    L.o_extract_features.restype = C.c_int
    L.o_extract_features.argtypes = [f32p, i32p, C.c_int, i32p, i32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                     f32p, i32p, i32p, i32p, i32p, i32p, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    od64 = _opt(f64p)
    L.o_icp_align.restype = C.c_int
    L.o_icp_align.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, f32p, C.POINTER(C.c_int),
                              C.POINTER(C.c_double), of32]
    L.o_icperr_create.restype = C.c_void_p
    L.o_icperr_create.argtypes = [f32p, C.c_int, f32p, C.c_int]
    L.o_icperr_destroy.argtypes = [C.c_void_p]
    L.o_icperr_evaluate.restype = C.c_double
    L.o_icperr_evaluate.argtypes = [C.c_void_p, f64p]
    L.o_icperr_yaw_search.restype = C.c_int
    L.o_icperr_yaw_search.argtypes = [C.c_void_p, f64p, f64p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.o_ndt_create.restype = C.c_void_p
    L.o_ndt_create.argtypes = [C.c_float, C.c_double, C.c_double, C.c_int]
    L.o_ndt_destroy.argtypes = [C.c_void_p]
    L.o_ndt_set_target.restype = C.c_int
    L.o_ndt_set_target.argtypes = [C.c_void_p, f32p, C.c_int]
    L.o_ndt_set_source.argtypes = [C.c_void_p, f32p, C.c_int]
    L.o_ndt_get_voxels.argtypes = [C.c_void_p, i32p, i32p, f32p, f64p, f64p]
    L.o_ndt_grid_geometry.argtypes = [C.c_void_p, i32p, i32p]
    L.o_ndt_derivatives.restype = C.c_double
    L.o_ndt_derivatives.argtypes = [C.c_void_p, f64p, f64p, od64, C.POINTER(C.c_longlong)]
    L.o_ndt_pose_to_matrix.argtypes = [f64p, f32p]
    L.o_ndt_align.argtypes = [C.c_void_p, f32p, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.o_ndt_fitness.restype = C.c_double
    L.o_ndt_fitness.argtypes = [C.c_void_p, f32p]
    L.o_svd_solve6.argtypes = [f64p, f64p, f64p]
    L.o_o3d_voxel_down_sample.restype = C.c_int
    L.o_o3d_voxel_down_sample.argtypes = [f64p, C.c_int, C.c_double, f64p, oi32]
    L.o_gicp_normals_covs.argtypes = [f64p, C.c_int, C.c_int, C.c_double, od64, od64, C.c_int]
    L.o_gicp_tree_create.restype = C.c_void_p
    L.o_gicp_tree_create.argtypes = [f64p, C.c_int]
    L.o_gicp_tree_destroy.argtypes = [C.c_void_p]
    L.o_gicp_linearize.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_int, C.c_void_p, f64p, C.c_double, f64p, oi32, C.c_int]
    L.o_gicp_solve_update.restype = C.c_int
    L.o_gicp_solve_update.argtypes = [f64p, f64p]
    L.o_gicp_register.restype = C.c_int
    L.o_gicp_register.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_int, f64p, C.c_double, C.c_double, C.c_double, C.c_int,
                                  f64p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    _LIB = L
    return L


def omp_threads(requested):
    """Size of the OpenMP team a `num_threads(requested)` region really gets in this process."""
    L = lib()
    L.o_omp_threads.restype = C.c_int
    L.o_omp_threads.argtypes = [C.c_int]
    return int(L.o_omp_threads(int(requested)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------- voxel grid
def voxel_grid(pts_xyzi, leaf, min_points=0):
    """pcl::VoxelGrid restatement. Returns dict(out, voxel_of_point, out_voxel_idx, refused)."""
    L = lib()
    p = _f32(pts_xyzi).reshape(-1, 4)
    n = p.shape[0]
    leaf3 = (leaf, leaf, leaf) if np.isscalar(leaf) else tuple(leaf)
    out = np.empty((max(n, 1), 4), np.float32)
    vop = np.empty(max(n, 1), np.int32)
    ovi = np.empty(max(n, 1), np.int32)
    refused = C.c_int(0)
    m = L.o_voxel_grid(p, n, leaf3[0], leaf3[1], leaf3[2], min_points, out, vop, C.byref(refused), ovi)
    return dict(out=out[:m].copy(), voxel_of_point=vop[:n].copy(), out_voxel_idx=ovi[:m].copy(), refused=bool(refused.value))


# ---------------------------------------------------------------- scan-to-map
class Scan2Map:
    def __init__(self, threads=1):
        self.L = lib()
        self.h = self.L.o_s2m_create(threads)
        self.nc = self.ns = 0

    def __del__(self):
        try:
            self.L.o_s2m_destroy(self.h)
        except Exception:
            pass

    def set_map(self, corner, surf):
        c, s = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        self.L.o_s2m_set_map(self.h, c, len(c), s, len(s))

    def set_scan(self, corner, surf):
        c, s = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        self.nc, self.ns = len(c), len(s)
        self.L.o_s2m_set_scan(self.h, c, len(c), s, len(s))

    def set_state(self, degenerate, matP=None):
        self.L.o_s2m_set_state(self.h, int(degenerate), None if matP is None else _f32(matP).reshape(36))

    def get_state(self):
        d = C.c_int(0)
        P = np.zeros(36, np.float32)
        self.L.o_s2m_get_state(self.h, C.byref(d), P)
        return bool(d.value), P.reshape(6, 6)

    def iterate(self, pose, it):
        pose = _f32(pose).copy()
        nsel, ran = C.c_int(0), C.c_int(0)
        AtA, AtB, X = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32)
        conv = self.L.o_s2m_iterate(self.h, pose, it, C.byref(nsel), C.byref(ran), AtA, AtB, X)
        return dict(pose=pose, converged=bool(conv), n_sel=nsel.value, ran=bool(ran.value),
                    AtA=AtA.reshape(6, 6), AtB=AtB, X=X)

    def get_pass(self, which):
        n = self.ns if which else self.nc
        idx = np.empty((n, 5), np.int32)
        d2 = np.empty((n, 5), np.float32)
        coeff = np.empty((n, 4), np.float32)
        flag = np.empty(n, np.uint8)
        self.L.o_s2m_get_pass(self.h, which, idx, d2, coeff, flag)
        return dict(idx=idx, d2=d2, coeff=coeff, flag=flag)

    def solve(self, pose, max_iters=30, edge_min=10, surf_min=100):
        pose = _f32(pose).copy()
        it, conv = C.c_int(0), C.c_int(0)
        hist = np.zeros((max_iters, 6), np.float32)
        nsel = np.zeros(max_iters, np.int32)
        rc = self.L.o_s2m_solve(self.h, pose, max_iters, edge_min, surf_min, C.byref(it), C.byref(conv), hist, nsel)
        return dict(rc=rc, pose=pose, iters=it.value, converged=bool(conv.value), pose_hist=hist[:it.value],
                    nsel_hist=nsel[:it.value])


def knn(pts, queries, k, brute=False, threads=8):
    L = lib()
    p, q = _f32(pts).reshape(-1, 4), _f32(queries).reshape(-1, 4)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    (L.o_knn_brute if brute else L.o_knn)(p, len(p), q, len(q), k, idx, d2, threads)
    return idx, d2


def transform_cloud(pts, pose, threads=1):
    L = lib()
    p = _f32(pts).reshape(-1, 4)
    out = np.empty_like(p)
    L.o_transform_cloud(p, len(p), _f32(pose), out, threads)
    return out


def pose_affine(pose):
    t = np.zeros(12, np.float32)
    lib().o_pose_affine(_f32(pose), t)
    return t.reshape(3, 4)


# ---------------------------------------------------------------- front end
def project(raw_xyzirt, N_SCAN, H, imu=None, t_cur=0.0, downsample=1, rmin=1.0, rmax=1000.0, deskew=True):
    """raw_xyzirt: structured/byte array of n*32 bytes. imu: (time, rx, ry, rz) float64 arrays or None."""
    L = lib()
    raw = np.ascontiguousarray(raw_xyzirt).view(np.uint8).reshape(-1)
    n = raw.size // 32
    cells = N_SCAN * H
    if imu is None:
        z = np.zeros(1, np.float64)
        imu = (z, z, z, z)
        n_imu = 0
    else:
        imu = tuple(np.ascontiguousarray(a, np.float64) for a in imu)
        n_imu = len(imu[0])
    rm = np.empty(cells, np.float32)
    fc = np.empty((cells, 4), np.float32)
    win = np.empty(cells, np.int32)
    ext = np.empty((cells, 4), np.float32)
    col = np.empty(cells, np.int32)
    rng = np.empty(cells, np.float32)
    sr = np.empty(N_SCAN, np.int32)
    er = np.empty(N_SCAN, np.int32)
    m = L.o_project(raw, n, N_SCAN, H, downsample, rmin, rmax, imu[0], imu[1], imu[2], imu[3], n_imu, t_cur,
                    1 if deskew else -1, rm, fc, win, ext, col, rng, sr, er)
    return dict(range_mat=rm.reshape(N_SCAN, H), full_cloud=fc, winner=win.reshape(N_SCAN, H), extracted=ext[:m].copy(),
                pointColInd=col[:m].copy(), pointRange=rng[:m].copy(), startRingIndex=sr, endRingIndex=er)


def imu_deskew_info(stamp, quat_xyzw, gyro, t_cur, t_end, capacity=2000):
    """imuDeskewInfo restatement (imageProjection.cpp:305-362). Returns dict(imu=(t, rx, ry, rz), imuAvailable, n_popped, rpy)."""
    stamp = np.ascontiguousarray(stamp, np.float64).reshape(-1)
    n = len(stamp)
    quat = np.ascontiguousarray(quat_xyzw if quat_xyzw is not None else np.tile([0.0, 0, 0, 1], (max(n, 1), 1)), np.float64).reshape(-1, 4)
    gyro = np.ascontiguousarray(gyro, np.float64).reshape(-1, 3)
    if n == 0:
        stamp, quat, gyro = np.zeros(1), np.zeros((1, 4)), np.zeros((1, 3))
    t, rx, ry, rz = (np.zeros(capacity, np.float64) for _ in range(4))
    nt, npop, avail = C.c_int(0), C.c_int(0), C.c_int(0)
    rpy = np.full(3, np.nan, np.float32)
    st = lib().o_imu_deskew_info(stamp, quat, gyro, n, float(t_cur), float(t_end), t, rx, ry, rz, capacity,
                                 C.byref(nt), C.byref(npop), C.byref(avail), rpy)
    if st != 0:
        raise RuntimeError("imu table overflow")
    k = nt.value
    return dict(imu=(t[:k].copy(), rx[:k].copy(), ry[:k].copy(), rz[:k].copy()), imuAvailable=bool(avail.value), n_popped=npop.value,
                rpy=None if np.isnan(rpy[0]) else rpy.copy())


def curvature_masks(pointRange, pointColInd):
    L = lib()
    r = _f32(pointRange)
    c = np.ascontiguousarray(pointColInd, np.int32)
    n = len(r)
    curv = np.empty(n, np.float32)
    picked = np.empty(n, np.int32)
    label = np.empty(n, np.int32)
    L.o_curvature_masks(r, c, n, curv, picked, label)
    return curv, picked, label


def extract_features(proj, edge_th=1.0, surf_th=0.1, surf_leaf=0.4, stable=True):
    """proj: dict from project(). Returns corner/surf clouds and the intermediate arrays."""
    L = lib()
    ext = proj["extracted"]
    n = len(ext)
    N_SCAN = len(proj["startRingIndex"])
    curv, picked, label = curvature_masks(proj["pointRange"], proj["pointColInd"])
    picked_after_mask = picked.copy()
    cidx = np.empty(N_SCAN * 6 * 20 + 1, np.int32)
    sidx = np.empty(max(n, 1), np.int32)
    srs = np.empty(N_SCAN + 1, np.int32)
    sds = np.empty((max(n, 1), 4), np.float32)
    ns, nds = C.c_int(0), C.c_int(0)
    nc = L.o_extract_features(np.ascontiguousarray(ext), proj["pointColInd"], n, proj["startRingIndex"], proj["endRingIndex"],
                              N_SCAN, edge_th, surf_th, surf_leaf, 1 if stable else 0, curv, picked, label,
                              cidx, sidx, srs, sds, C.byref(ns), C.byref(nds))
    return dict(curvature=curv, picked_mask=picked_after_mask, picked=picked, label=label,
                corner_idx=cidx[:nc].copy(), corner=ext[cidx[:nc]].copy(), surf_idx=sidx[:ns.value].copy(),
                surf_ring_start=srs, surf=sds[:nds.value].copy())


# ---------------------------------------------------------------- Open3D GICP path (Multi_LiCa)
def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def o3d_voxel_down_sample(pts, voxel):
    """Open3D voxel_down_sample restatement; output in ascending (z, y, x) voxel order. Returns (points, rank_of_point)."""
    p = _f64(pts).reshape(-1, 3)
    out = np.empty((max(len(p), 1), 3), np.float64)
    rank = np.empty(max(len(p), 1), np.int32)
    m = lib().o_o3d_voxel_down_sample(p, len(p), float(voxel), out, rank)
    return out[:m].copy(), rank[:len(p)].copy()


def gicp_normals_covs(pts, knn=30, eps=0.005, threads=8):
    p = _f64(pts).reshape(-1, 3)
    nrm = np.empty((len(p), 3), np.float64)
    cov = np.empty((len(p), 9), np.float64)
    lib().o_gicp_normals_covs(p, len(p), int(knn), float(eps), nrm, cov, threads)
    return nrm, cov.reshape(-1, 3, 3)


class GicpOracle:
    """One target (with its kd-tree) and one source, covariances given; linearize / register as Open3D does."""

    def __init__(self, src, src_cov, tgt, tgt_cov, threads=8):
        self.L = lib()
        self.src, self.src_cov = _f64(src).reshape(-1, 3), _f64(src_cov).reshape(-1, 9)
        self.tgt, self.tgt_cov = _f64(tgt).reshape(-1, 3), _f64(tgt_cov).reshape(-1, 9)
        self.threads = threads
        self.tree = self.L.o_gicp_tree_create(self.tgt, len(self.tgt))

    def __del__(self):
        if getattr(self, "tree", None):
            self.L.o_gicp_tree_destroy(self.tree)
            self.tree = None

    def linearize(self, T, max_corr, want_corr=False, begin=0, end=None):
        """sums[30] (+ correspondences) of source[begin:end] at transform T."""
        end = len(self.src) if end is None else end
        s, c = self.src[begin:end], self.src_cov[begin:end]
        sums = np.zeros(30, np.float64)
        corr = np.empty(max(len(s), 1), np.int32) if want_corr else None
        self.L.o_gicp_linearize(np.ascontiguousarray(s), np.ascontiguousarray(c), len(s), self.tgt, self.tgt_cov, len(self.tgt),
                                self.tree, _f64(T).reshape(16), float(max_corr), sums, corr, self.threads)
        return (sums, corr[:len(s)]) if want_corr else sums

    def solve_update(self, sums):
        U = np.empty(16, np.float64)
        ok = self.L.o_gicp_solve_update(_f64(sums), U)
        return U.reshape(4, 4), bool(ok)

    def register(self, init, max_corr, rel_fit, rel_rmse, max_it):
        T = np.empty(16, np.float64)
        fit, rmse = C.c_double(), C.c_double()
        it = self.L.o_gicp_register(self.src, self.src_cov, len(self.src), self.tgt, self.tgt_cov, len(self.tgt),
                                    _f64(init).reshape(16), float(max_corr), float(rel_fit), float(rel_rmse), int(max_it),
                                    T, C.byref(fit), C.byref(rmse), self.threads)
        return dict(transformation=T.reshape(4, 4), fitness=fit.value, inlier_rmse=rmse.value, iterations=it)


# ---------------------------------------------------------------- PCL NDT path (multi_lidar calibrator)
class NdtOracle:
    """pcl::NormalDistributionsTransform<PointXYZ, PointXYZ> restatement (serial, like PCL's)."""

    def __init__(self, resolution=1.0, step_size=0.1, epsilon=0.01, max_iterations=400):
        self.L = lib()
        self.h = self.L.o_ndt_create(float(resolution), float(step_size), float(epsilon), int(max_iterations))
        self.n_voxels = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.o_ndt_destroy(self.h)
            self.h = None

    def set_target(self, xyz):
        p = _f32(xyz).reshape(-1, 3)
        self.n_voxels = self.L.o_ndt_set_target(self.h, p, len(p))
        return self.n_voxels

    def set_source(self, xyz):
        p = _f32(xyz).reshape(-1, 3)
        self.L.o_ndt_set_source(self.h, p, len(p))

    def voxels(self):
        m = max(self.n_voxels, 1)
        idx, npts = np.empty(m, np.int32), np.empty(m, np.int32)
        cen, mean, icov = np.empty((m, 3), np.float32), np.empty((m, 3), np.float64), np.empty((m, 9), np.float64)
        self.L.o_ndt_get_voxels(self.h, idx, npts, cen, mean, icov)
        k = self.n_voxels
        mn, dv = np.empty(3, np.int32), np.empty(3, np.int32)
        self.L.o_ndt_grid_geometry(self.h, mn, dv)
        return dict(index=idx[:k], npts=npts[:k], centroid=cen[:k], mean=mean[:k], icov=icov[:k].reshape(-1, 3, 3), min_b=mn, div_b=dv)

    def derivatives(self, p):
        g, H, pairs = np.zeros(6), np.zeros(36), C.c_longlong()
        s = self.L.o_ndt_derivatives(self.h, _f64(p), g, H, C.byref(pairs))
        return s, g, H.reshape(6, 6), pairs.value

    def align(self, guess):
        T = np.empty(16, np.float32)
        it, conv, prob, ev = C.c_int(), C.c_int(), C.c_double(), C.c_int()
        self.L.o_ndt_align(self.h, _f32(guess).reshape(16), T, C.byref(it), C.byref(conv), C.byref(prob), C.byref(ev))
        return dict(transformation=T.reshape(4, 4), iterations=it.value, converged=bool(conv.value),
                    transformation_probability=prob.value, evaluations=ev.value)

    def fitness(self, T):
        return self.L.o_ndt_fitness(self.h, _f32(T).reshape(16))


def ndt_pose_to_matrix(p):
    T = np.empty(16, np.float32)
    lib().o_ndt_pose_to_matrix(_f64(p), T)
    return T.reshape(4, 4)


def svd_solve6(H, b):
    x = np.empty(6)
    lib().o_svd_solve6(_f64(H).reshape(36), _f64(b), x)
    return x


# ---------------------------------------------------------------- SensorsCalibration yaw grid search
class IcpErrorOracle:
    def __init__(self, tgt, src):
        self.L = lib()
        t, s_ = _f32(tgt).reshape(-1, 3), _f32(src).reshape(-1, 3)
        self.h = self.L.o_icperr_create(t, len(t), s_, len(s_))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.o_icperr_destroy(self.h)
            self.h = None

    def evaluate(self, T):
        return self.L.o_icperr_evaluate(self.h, _f64(T).reshape(16))

    def yaw_search(self, init):
        T = np.empty(16, np.float64)
        yaw, err = C.c_double(), C.c_double()
        ev = self.L.o_icperr_yaw_search(self.h, _f64(init).reshape(16), T, C.byref(yaw), C.byref(err))
        return dict(transform=T.reshape(4, 4), best_yaw=yaw.value, min_error=err.value, evaluations=ev)


def icp_align(src, tgt, max_corr_dist=30.0, max_iterations=100, transformation_epsilon=1e-6, euclidean_fitness_epsilon=1e-6):
    """pcl::IterativeClosestPoint restatement (identity guess), as mapOptmization.cpp:559-586 configures it."""
    s_, t_ = _f32(src).reshape(-1, 3), _f32(tgt).reshape(-1, 3)
    T = np.empty(16, np.float32)
    hist = np.zeros((max(max_iterations, 1), 16), np.float32)
    conv, fit = C.c_int(), C.c_double()
    it = lib().o_icp_align(s_, len(s_), t_, len(t_), float(max_corr_dist), int(max_iterations), float(transformation_epsilon),
                           float(euclidean_fitness_epsilon), T, C.byref(conv), C.byref(fit), hist)
    return dict(transformation=T.reshape(4, 4), iterations=it, converged=bool(conv.value), fitness_score=fit.value,
                history=hist[:it].reshape(-1, 4, 4))


# ---------------------------------------------------------------- PointClouds_Fusion front end (numpy restatement)
def fuse_clouds(clouds, transforms, external_bounds=None, internal_bounds=None):
    """fusion_pointclouds.cpp:55-115 — pcl::transformPointCloud with a double matrix (sums in double, stored float), `+`
    concatenation in the given order, pcl::PassThrough on x, z, y (float limits, keeps min <= v <= max, drops non-finite
    points), pcl::ConditionalRemoval with a ConditionOr of GT / LT comparisons in double (keeps what lies outside the box)."""
    parts = []
    for c, T in zip(clouds, transforms):
        c = _f32(c).reshape(-1, c.shape[1])[:, :4]
        if T is None:
            parts.append(c.copy())
            continue
        T = _f64(T)
        x, y, z = (c[:, k].astype(np.float64) for k in range(3))
        out = c.copy()
        for r in range(3):
            out[:, r] = (T[r, 0] * x + T[r, 1] * y + T[r, 2] * z + T[r, 3]).astype(np.float32)
        parts.append(out)
    fused = np.concatenate(parts) if parts else np.zeros((0, 4), np.float32)
    keep = np.ones(len(fused), bool)
    if external_bounds is not None:
        lo, hi = np.asarray(external_bounds[0], np.float64).astype(np.float32), np.asarray(external_bounds[1], np.float64).astype(np.float32)
        keep &= np.isfinite(fused[:, :3]).all(1)
        for d in range(3):
            keep &= ~((fused[:, d] < lo[d]) | (fused[:, d] > hi[d]))
    if internal_bounds is not None:
        lo, hi = np.asarray(internal_bounds[0], np.float64), np.asarray(internal_bounds[1], np.float64)
        v = fused[:, :3].astype(np.float64)
        keep &= ((v > hi) | (v < lo)).any(1)
    return fused[keep]


# ---- lio_sam/cloud_info on the wire — ROS1 serialisation written out with struct, independent of the
# C walk in csrc/b2_scan.cu. Follows msg/cloud_info.msg:1-35, sensor_msgs/PointCloud2 + PointField + std_msgs/Header, and
# what pcl::toROSMsg produces for a pcl::PointXYZI cloud (utility.h:286-295). TEST INFRASTRUCTURE ONLY.
def _ros_header(seq, stamp, frame_id):
    import struct
    f = frame_id.encode()
    return struct.pack("<IIII", seq, stamp[0], stamp[1], len(f)) + f


def _ros_cloud_xyzi(points_xyzi, stamp, frame_id):
    import struct
    p = np.ascontiguousarray(points_xyzi, np.float32).reshape(-1, 4)
    n = len(p)
    rec = np.zeros((n, 8), np.float32)
    rec[:, :3] = p[:, :3]; rec[:, 3] = 1.0; rec[:, 4] = p[:, 3]        # PointXYZI: data[3] = 1.0f, intensity at byte 16
    out = _ros_header(0, stamp, frame_id) + struct.pack("<II", 1, n) + struct.pack("<I", 4)
    for name, off in (("x", 0), ("y", 4), ("z", 8), ("intensity", 16)):
        out += struct.pack("<I", len(name)) + name.encode() + struct.pack("<IBI", off, 7, 1)
    out += struct.pack("<BII", 0, 32, 32 * n) + struct.pack("<I", 32 * n) + rec.tobytes() + struct.pack("<B", 1)
    return out


def _ros_cloud_empty():
    import struct
    return _ros_header(0, (0, 0), "") + struct.pack("<II", 0, 0) + struct.pack("<I", 0) + struct.pack("<BII", 0, 0, 0) + struct.pack("<I", 0) + b"\0"


def serialize_cloud_info(stage, seq=0, stamp=(0, 0), frame_id="", lidarFrame="", startRingIndex=(), endRingIndex=(), pointColInd=(),
                         pointRange=(), imuAvailable=0, odomAvailable=0, imuRollInit=0.0, imuPitchInit=0.0, imuYawInit=0.0,
                         initialGuess=(0.0,) * 6, cloud_deskewed=None, cloud_corner=None, cloud_surface=None):
    """stage 0: imageProjection.cpp:600-605 (arrays + cloud_deskewed); stage 1: featureExtraction.cpp:240-258 (arrays cleared,
    cloud_corner / cloud_surface added)."""
    import struct

    def arr(a, dt):
        a = np.ascontiguousarray(a, dt).reshape(-1) if stage == 0 else np.zeros(0, dt)
        return struct.pack("<I", len(a)) + a.tobytes()

    out = _ros_header(seq, stamp, frame_id)
    out += arr(startRingIndex, np.int32) + arr(endRingIndex, np.int32) + arr(pointColInd, np.int32) + arr(pointRange, np.float32)
    out += struct.pack("<qq", imuAvailable, odomAvailable)
    out += np.asarray([imuRollInit, imuPitchInit, imuYawInit, *initialGuess], np.float32).tobytes()
    out += _ros_cloud_xyzi(cloud_deskewed, stamp, lidarFrame)
    if stage == 1:
        out += _ros_cloud_xyzi(cloud_corner, stamp, lidarFrame) + _ros_cloud_xyzi(cloud_surface, stamp, lidarFrame)
    else:
        out += _ros_cloud_empty() * 2
    out += _ros_cloud_empty() * 4
    return out
