"""ctypes binding of libb2reg.so (include/b2reg.h). No fallbacks: if the CUDA library is missing or a call fails,
this raises — the product never routes around the device path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2_LIB") or os.path.join(HERE, "libb2reg.so")   # B2_LIB: kernel experiments only
_LIB = None


class B2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb2reg error {code}: {msg}")
        self.code = code


class ScanParams(C.Structure):
    _fields_ = [("n_scan", C.c_int), ("horizon_scan", C.c_int), ("downsample_rate", C.c_int), ("lidar_min_range", C.c_float),
                ("lidar_max_range", C.c_float), ("edge_threshold", C.c_float), ("surf_threshold", C.c_float),
                ("odometry_surf_leaf_size", C.c_float)]


class S2MParams(C.Structure):
    _fields_ = [("edge_feature_min_valid_num", C.c_int), ("surf_feature_min_valid_num", C.c_int),
                ("max_iterations", C.c_int), ("min_correspondences", C.c_int), ("knn_max_dist", C.c_float),
                ("degenerate_eigen_threshold", C.c_float), ("max_batch", C.c_int)]


class CloudInfoMeta(C.Structure):
    _fields_ = [("seq", C.c_uint32), ("stamp_sec", C.c_uint32), ("stamp_nsec", C.c_uint32), ("frame_id", C.c_char_p),
                ("cloud_frame_id", C.c_char_p), ("imu_available", C.c_int64), ("odom_available", C.c_int64),
                ("imu_roll_init", C.c_float), ("imu_pitch_init", C.c_float), ("imu_yaw_init", C.c_float),
                ("initial_guess_x", C.c_float), ("initial_guess_y", C.c_float), ("initial_guess_z", C.c_float),
                ("initial_guess_roll", C.c_float), ("initial_guess_pitch", C.c_float), ("initial_guess_yaw", C.c_float)]


class Cloud2View(C.Structure):
    _fields_ = [("data", C.c_void_p), ("height", C.c_uint32), ("width", C.c_uint32), ("point_step", C.c_uint32),
                ("row_step", C.c_uint32), ("n_fields", C.c_uint32), ("off_x", C.c_int32), ("off_y", C.c_int32),
                ("off_z", C.c_int32), ("off_intensity", C.c_int32), ("is_bigendian", C.c_uint8), ("is_dense", C.c_uint8)]


class CloudInfoView(C.Structure):
    _fields_ = [("seq", C.c_uint32), ("stamp_sec", C.c_uint32), ("stamp_nsec", C.c_uint32),
                ("frame_id", C.c_void_p), ("frame_id_len", C.c_uint32),
                ("start_ring_index", C.c_void_p), ("n_start_ring_index", C.c_uint32),
                ("end_ring_index", C.c_void_p), ("n_end_ring_index", C.c_uint32),
                ("point_col_ind", C.c_void_p), ("n_point_col_ind", C.c_uint32),
                ("point_range", C.c_void_p), ("n_point_range", C.c_uint32),
                ("imu_available", C.c_int64), ("odom_available", C.c_int64),
                ("imu_roll_init", C.c_float), ("imu_pitch_init", C.c_float), ("imu_yaw_init", C.c_float),
                ("initial_guess_x", C.c_float), ("initial_guess_y", C.c_float), ("initial_guess_z", C.c_float),
                ("initial_guess_roll", C.c_float), ("initial_guess_pitch", C.c_float), ("initial_guess_yaw", C.c_float),
                ("cloud_deskewed", Cloud2View), ("cloud_corner", Cloud2View), ("cloud_surface", Cloud2View),
                ("key_frame_cloud", Cloud2View), ("key_frame_color", Cloud2View), ("key_frame_poses", Cloud2View),
                ("key_frame_map", Cloud2View)]


class GicpParams(C.Structure):
    _fields_ = [("max_correspondence_distance", C.c_double), ("epsilon", C.c_double), ("relative_fitness", C.c_double),
                ("relative_rmse", C.c_double), ("max_iteration", C.c_int)]


# every symbol include/b2reg.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "b2_version", "b2_last_error", "b2_device_count", "b2_set_device", "b2_kernel_launch_count", "b2_trim_memory",
    "b2_voxel_create", "b2_voxel_destroy", "b2_voxel_set_leaf_size", "b2_voxel_set_min_points_per_voxel", "b2_voxel_filter", "b2_voxel_last_gpu_ms",
    "b2_knn_create", "b2_knn_destroy", "b2_knn_set_input_cloud", "b2_knn_nearest_k_search",
    "b2_s2m_default_params", "b2_s2m_create", "b2_s2m_destroy", "b2_s2m_set_map", "b2_s2m_set_scan", "b2_s2m_iterate",
    "b2_s2m_solve", "b2_s2m_set_state", "b2_s2m_get_pass", "b2_s2m_get_normal_equations", "b2_s2m_set_scan_batch",
    "b2_s2m_solve_batch", "b2_s2m_last_gpu_ms", "b2_s2m_rebuild_map_index", "b2_s2m_last_step_gpu_ms", "b2_s2m_count_candidates", "b2_transform_cloud",
    "b2_scan_default_params", "b2_scan_create", "b2_scan_destroy", "b2_scan_project", "b2_scan_extract_features",
    "b2_scan_last_gpu_ms", "b2_imu_deskew_info",
    "b2_cloud_info_parse", "b2_scan_write_cloud_info", "b2_scan_set_from_cloud_info", "b2_s2m_set_scan_downsampled",
    "b2_s2m_set_scan_from_front_end", "b2_s2m_get_scan",
    "b2_cloud_create", "b2_cloud_destroy", "b2_cloud_set_points", "b2_cloud_set_points_f32", "b2_cloud_size",
    "b2_cloud_get_points", "b2_cloud_get_normals", "b2_cloud_set_normals", "b2_cloud_voxel_down_sample",
    "b2_cloud_estimate_normals", "b2_cloud_transform", "b2_cloud_last_gpu_ms", "b2_cloud_set_points_sharded", "b2_cloud_estimate_normals_sharded",
    "b2_comm_unique_id", "b2_comm_create", "b2_comm_destroy", "b2_comm_rank", "b2_comm_allreduce_f64",
    "b2_gicp_default_params", "b2_gicp_create", "b2_gicp_destroy", "b2_gicp_set_params", "b2_gicp_set_target",
    "b2_gicp_set_source", "b2_gicp_set_source_slice", "b2_gicp_set_source_blocks", "b2_gicp_set_shard", "b2_gicp_peer_handle", "b2_gicp_set_peers", "b2_gicp_exchange_setup", "b2_gicp_linearize", "b2_gicp_align", "b2_gicp_get_history",
    "b2_gicp_last_gpu_ms", "b2_gicp_index_info", "b2_gicp_get_evaluation_ms",
    "b2_localmap_create", "b2_localmap_destroy", "b2_localmap_add_keyframe", "b2_localmap_num_keyframes", "b2_localmap_set_pose",
    "b2_localmap_clear_cache", "b2_localmap_extract", "b2_localmap_get", "b2_localmap_last_gpu_ms", "b2_s2m_set_map_from_localmap",
    "b2_nnerr_create", "b2_nnerr_destroy", "b2_nnerr_set_target", "b2_nnerr_set_source", "b2_nnerr_evaluate", "b2_nnerr_yaw_search",
    "b2_nnerr_last_gpu_ms",
    "b2_icp_create", "b2_icp_destroy", "b2_icp_set_max_correspondence_distance", "b2_icp_set_maximum_iterations",
    "b2_icp_set_transformation_epsilon", "b2_icp_set_euclidean_fitness_epsilon", "b2_icp_set_ransac_iterations", "b2_icp_set_input_source",
    "b2_icp_set_input_target", "b2_icp_align", "b2_icp_has_converged", "b2_icp_get_fitness_score", "b2_icp_get_final_transformation",
    "b2_icp_get_final_num_iteration", "b2_icp_last_gpu_ms",
    "b2_fusion_create", "b2_fusion_destroy", "b2_fusion_clear", "b2_fusion_add_cloud", "b2_fusion_set_external_bounds",
    "b2_fusion_set_internal_bounds", "b2_fusion_run", "b2_fusion_get", "b2_fusion_device_cloud", "b2_fusion_last_gpu_ms",
    "b2_ndt_create", "b2_ndt_destroy", "b2_ndt_set_transformation_epsilon", "b2_ndt_set_step_size", "b2_ndt_set_resolution",
    "b2_ndt_set_maximum_iterations", "b2_ndt_set_input_target", "b2_ndt_set_input_source", "b2_ndt_align", "b2_ndt_has_converged",
    "b2_ndt_get_final_transformation", "b2_ndt_get_fitness_score", "b2_ndt_get_transformation_probability",
    "b2_ndt_get_final_num_iteration", "b2_ndt_get_voxels", "b2_ndt_derivatives", "b2_ndt_last_gpu_ms",
]


class _Missing:
    def __init__(self, name):
        self.__dict__["_name"] = name

    def __call__(self, *a):
        raise ImportError(f"{self._name} is not exported by {LIB_PATH}")


class _ExperimentLib:
    """B2_LIB points at a kernel-experiment build that may predate an entry point: binding it must not fail, calling it does."""

    def __init__(self, cdll):
        self.__dict__["_c"] = cdll

    def __getattr__(self, name):
        try:
            return getattr(self._c, name)
        except AttributeError:
            m = _Missing(name)
            self.__dict__[name] = m
            return m


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m multi_sensor_slam_tookit_b200.build` "
                          "(there is no CPU fallback for this path)")
    L = C.CDLL(LIB_PATH)
    if os.environ.get("B2_LIB"):
        L = _ExperimentLib(L)
    vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    pi, pf = C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.b2_last_error.restype = C.c_char_p
    L.b2_set_device.argtypes = [i32]
    L.b2_voxel_create.argtypes = [C.POINTER(vp)]
    L.b2_voxel_destroy.argtypes = [vp]
    L.b2_voxel_set_leaf_size.argtypes = [vp, f32, f32, f32]
    L.b2_voxel_set_min_points_per_voxel.argtypes = [vp, C.c_uint]
    L.b2_voxel_last_gpu_ms.argtypes = [vp, pf]
    L.b2_voxel_filter.argtypes = [vp, vp, sz, sz, i32, vp, sz, sz, C.POINTER(sz), pi, vp]
    L.b2_knn_create.argtypes = [C.POINTER(vp), f32]
    L.b2_knn_destroy.argtypes = [vp]
    L.b2_knn_set_input_cloud.argtypes = [vp, vp, sz, sz]
    L.b2_knn_nearest_k_search.argtypes = [vp, vp, sz, sz, i32, vp, vp]
    L.b2_s2m_default_params.argtypes = [C.POINTER(S2MParams)]
    L.b2_s2m_default_params.restype = None
    L.b2_s2m_create.argtypes = [C.POINTER(vp), C.POINTER(S2MParams)]
    L.b2_s2m_destroy.argtypes = [vp]
    L.b2_s2m_set_map.argtypes = [vp, vp, sz, sz, vp, sz, sz]
    L.b2_s2m_set_scan.argtypes = [vp, vp, sz, sz, vp, sz, sz]
    L.b2_s2m_iterate.argtypes = [vp, vp, i32, pi, pi, pi, pi, vp]
    L.b2_s2m_solve.argtypes = [vp, vp, i32, pi, pi, pi, vp, pi, vp]
    L.b2_s2m_set_state.argtypes = [vp, i32, vp]
    L.b2_s2m_get_pass.argtypes = [vp, i32, vp, vp, vp, vp]
    L.b2_s2m_get_normal_equations.argtypes = [vp, vp, vp, vp]
    L.b2_s2m_set_scan_batch.argtypes = [vp, i32, vp, sz, vp, vp, sz, vp]
    L.b2_s2m_solve_batch.argtypes = [vp, vp, i32, vp, vp, vp]
    L.b2_s2m_last_gpu_ms.argtypes = [vp, pf, pi]
    L.b2_s2m_rebuild_map_index.argtypes = [vp]
    L.b2_s2m_count_candidates.argtypes = [vp, i32, C.POINTER(C.c_ulonglong)]
    L.b2_s2m_last_step_gpu_ms.argtypes = [vp, pf]
    L.b2_transform_cloud.argtypes = [vp, sz, sz, vp, vp, sz]
    L.b2_scan_default_params.argtypes = [C.POINTER(ScanParams)]
    L.b2_scan_default_params.restype = None
    L.b2_scan_create.argtypes = [C.POINTER(vp), C.POINTER(ScanParams)]
    L.b2_scan_destroy.argtypes = [vp]
    L.b2_scan_project.argtypes = [vp, vp, sz, vp, vp, vp, vp, i32, C.c_double, i32, C.POINTER(sz), vp, vp, vp, vp, vp, vp, vp]
    L.b2_scan_extract_features.argtypes = [vp, C.POINTER(sz), vp, vp, C.POINTER(sz), vp, vp, vp, vp]
    L.b2_scan_last_gpu_ms.argtypes = [vp, pf]
    L.b2_imu_deskew_info.argtypes = [vp, vp, vp, i32, C.c_double, C.c_double, vp, vp, vp, vp, i32, pi, pi, pi, vp]
    L.b2_cloud_info_parse.argtypes = [vp, sz, C.POINTER(CloudInfoView)]
    L.b2_scan_write_cloud_info.argtypes = [vp, C.POINTER(CloudInfoMeta), i32, vp, sz, C.POINTER(sz)]
    L.b2_scan_set_from_cloud_info.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.b2_s2m_set_scan_downsampled.argtypes = [vp, vp, vp, sz, sz, vp, vp, sz, sz, C.POINTER(sz), C.POINTER(sz)]
    L.b2_s2m_set_scan_from_front_end.argtypes = [vp, vp, vp, vp, C.POINTER(sz), C.POINTER(sz)]
    L.b2_s2m_get_scan.argtypes = [vp, i32, vp, sz, C.POINTER(sz)]
    pd, dbl = C.POINTER(C.c_double), C.c_double
    L.b2_cloud_create.argtypes = [C.POINTER(vp)]
    L.b2_cloud_destroy.argtypes = [vp]
    L.b2_cloud_set_points.argtypes = [vp, vp, sz]
    L.b2_cloud_set_points_f32.argtypes = [vp, vp, sz, sz]
    L.b2_cloud_size.argtypes = [vp, C.POINTER(sz), pi]
    L.b2_cloud_get_points.argtypes = [vp, vp]
    L.b2_cloud_get_normals.argtypes = [vp, vp]
    L.b2_cloud_set_normals.argtypes = [vp, vp]
    L.b2_cloud_voxel_down_sample.argtypes = [vp, dbl, C.POINTER(vp), vp]
    L.b2_cloud_estimate_normals.argtypes = [vp, i32]
    L.b2_cloud_estimate_normals_sharded.argtypes = [vp, i32, vp]
    L.b2_cloud_set_points_sharded.argtypes = [vp, vp, sz, vp]
    L.b2_cloud_transform.argtypes = [vp, vp]
    L.b2_cloud_last_gpu_ms.argtypes = [vp, pf]
    L.b2_comm_unique_id.argtypes = [vp]
    L.b2_comm_create.argtypes = [C.POINTER(vp), vp, i32, i32]
    L.b2_comm_destroy.argtypes = [vp]
    L.b2_comm_rank.argtypes = [vp, pi, pi]
    L.b2_comm_allreduce_f64.argtypes = [vp, vp, sz]
    L.b2_gicp_default_params.argtypes = [C.POINTER(GicpParams)]
    L.b2_gicp_default_params.restype = None
    L.b2_gicp_create.argtypes = [C.POINTER(vp), C.POINTER(GicpParams)]
    L.b2_gicp_destroy.argtypes = [vp]
    L.b2_gicp_set_params.argtypes = [vp, C.POINTER(GicpParams)]
    L.b2_gicp_set_target.argtypes = [vp, vp]
    L.b2_gicp_set_source.argtypes = [vp, vp]
    L.b2_gicp_set_source_slice.argtypes = [vp, vp, sz, sz]
    L.b2_gicp_set_source_blocks.argtypes = [vp, vp, i32, i32]
    L.b2_gicp_peer_handle.argtypes = [vp, vp]
    L.b2_gicp_set_peers.argtypes = [vp, i32, i32, vp]
    L.b2_gicp_exchange_setup.argtypes = [vp]
    L.b2_gicp_set_shard.argtypes = [vp, i32, i32, vp]
    L.b2_gicp_linearize.argtypes = [vp, vp, vp, vp]
    L.b2_gicp_align.argtypes = [vp, vp, vp, pd, pd, pi, pi]
    L.b2_gicp_get_history.argtypes = [vp, vp, vp, i32, pi]
    L.b2_gicp_last_gpu_ms.argtypes = [vp, pf, pi]
    L.b2_gicp_get_evaluation_ms.argtypes = [vp, vp, i32, pi]
    L.b2_gicp_index_info.argtypes = [vp, pd, pd, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.b2_localmap_create.argtypes = [C.POINTER(vp), f32, f32]
    L.b2_localmap_destroy.argtypes = [vp]
    L.b2_localmap_add_keyframe.argtypes = [vp, vp, sz, sz, vp, sz, sz, vp, pi]
    L.b2_localmap_num_keyframes.argtypes = [vp, pi]
    L.b2_localmap_set_pose.argtypes = [vp, i32, vp]
    L.b2_localmap_clear_cache.argtypes = [vp]
    L.b2_localmap_extract.argtypes = [vp, vp, i32, C.POINTER(sz), C.POINTER(sz)]
    L.b2_localmap_get.argtypes = [vp, i32, vp, sz, sz, C.POINTER(sz)]
    L.b2_localmap_last_gpu_ms.argtypes = [vp, pf, C.POINTER(sz)]
    L.b2_s2m_set_map_from_localmap.argtypes = [vp, vp]
    L.b2_nnerr_create.argtypes = [C.POINTER(vp)]
    L.b2_nnerr_destroy.argtypes = [vp]
    L.b2_nnerr_set_target.argtypes = [vp, vp, sz, sz]
    L.b2_nnerr_set_source.argtypes = [vp, vp, sz, sz]
    L.b2_nnerr_evaluate.argtypes = [vp, vp, pd, C.POINTER(sz)]
    L.b2_nnerr_yaw_search.argtypes = [vp, vp, vp, pd, pd, pi]
    L.b2_nnerr_last_gpu_ms.argtypes = [vp, pf]
    L.b2_icp_create.argtypes = [C.POINTER(vp)]
    L.b2_icp_destroy.argtypes = [vp]
    L.b2_icp_set_max_correspondence_distance.argtypes = [vp, dbl]
    L.b2_icp_set_maximum_iterations.argtypes = [vp, i32]
    L.b2_icp_set_transformation_epsilon.argtypes = [vp, dbl]
    L.b2_icp_set_euclidean_fitness_epsilon.argtypes = [vp, dbl]
    L.b2_icp_set_ransac_iterations.argtypes = [vp, i32]
    L.b2_icp_set_input_source.argtypes = [vp, vp, sz, sz]
    L.b2_icp_set_input_target.argtypes = [vp, vp, sz, sz]
    L.b2_icp_align.argtypes = [vp, vp, vp, sz]
    L.b2_icp_has_converged.argtypes = [vp, pi]
    L.b2_icp_get_fitness_score.argtypes = [vp, pd]
    L.b2_icp_get_final_transformation.argtypes = [vp, vp]
    L.b2_icp_get_final_num_iteration.argtypes = [vp, pi]
    L.b2_icp_last_gpu_ms.argtypes = [vp, pf, pi]
    L.b2_fusion_create.argtypes = [C.POINTER(vp)]
    L.b2_fusion_destroy.argtypes = [vp]
    L.b2_fusion_clear.argtypes = [vp]
    L.b2_fusion_add_cloud.argtypes = [vp, vp, sz, sz, vp]
    L.b2_fusion_set_external_bounds.argtypes = [vp, i32, vp, vp]
    L.b2_fusion_set_internal_bounds.argtypes = [vp, i32, vp, vp]
    L.b2_fusion_run.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.b2_fusion_get.argtypes = [vp, vp, sz, sz, C.POINTER(sz)]
    L.b2_fusion_device_cloud.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.b2_fusion_last_gpu_ms.argtypes = [vp, pf]
    L.b2_ndt_create.argtypes = [C.POINTER(vp)]
    L.b2_ndt_destroy.argtypes = [vp]
    L.b2_ndt_set_transformation_epsilon.argtypes = [vp, dbl]
    L.b2_ndt_set_step_size.argtypes = [vp, dbl]
    L.b2_ndt_set_resolution.argtypes = [vp, f32]
    L.b2_ndt_set_maximum_iterations.argtypes = [vp, i32]
    L.b2_ndt_set_input_target.argtypes = [vp, vp, sz, sz]
    L.b2_ndt_set_input_source.argtypes = [vp, vp, sz, sz]
    L.b2_ndt_align.argtypes = [vp, vp, vp, sz]
    L.b2_ndt_has_converged.argtypes = [vp, pi]
    L.b2_ndt_get_final_transformation.argtypes = [vp, vp]
    L.b2_ndt_get_fitness_score.argtypes = [vp, pd]
    L.b2_ndt_get_transformation_probability.argtypes = [vp, pd]
    L.b2_ndt_get_final_num_iteration.argtypes = [vp, pi]
    L.b2_ndt_get_voxels.argtypes = [vp, sz, C.POINTER(sz), vp, vp, vp, vp, vp, vp, vp]
    L.b2_ndt_derivatives.argtypes = [vp, vp, pd, vp, vp, C.POINTER(C.c_longlong)]
    L.b2_ndt_last_gpu_ms.argtypes = [vp, pf, pi, pi, C.POINTER(C.c_longlong)]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("b2_last_error", "b2_s2m_default_params", "b2_scan_default_params", "b2_kernel_launch_count",
                        "b2_gicp_default_params"):
            fn.restype = C.c_int
    L.b2_kernel_launch_count.restype = C.c_ulonglong
    _LIB = L
    return L


def check(code):
    if code != 0:
        raise B2Error(code, lib().b2_last_error().decode(errors="replace"))


def ptr(a):
    """Address of a numpy array's buffer (or None). ndarray.ctypes.data_as costs ~5 us per call — a seventh of the C1
    end-to-end step went into it; the buffer protocol gives the same address in under 1 us."""
    if a is None:
        return None
    try:
        return C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError, BufferError):          # read-only or empty arrays
        return a.ctypes.data


def as_points(a, min_cols=3):
    """C-contiguous float32 (n, c) view; returns (array, stride_bytes)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < min_cols:
        raise ValueError(f"expected (n, >={min_cols}) float32 points, got {a.shape}")
    return a, a.shape[1] * 4          # (numpy reports stride 0 for empty arrays)
