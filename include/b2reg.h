/* b2reg.h — C ABI of libb2reg.so: the B200-native (sm_100a) scan-registration hot path.
 *
 * The reference (JBaien/multi-sensor-slam-tookit) has no plugin/FFI layer; its seams for this path are
 * PCL class interfaces, four mapOptimization member functions and Open3D Python calls.
 * Every entry point below names the reference interface (file:line under /root/reference) it replaces.
 * INTEGRATION.md shows the reference-side binding a maintainer would add for each.
 *
 * Conventions
 *   - plain C, no exceptions, no callbacks, no torch/CUDA types in signatures;
 *   - every point-array argument is (base, stride_bytes, n): the first 3 floats at each stride are x,y,z and,
 *     where the op carries intensity, the float at byte offset `B2_INTENSITY_OFFSET(stride)`:
 *     PCL's 32 B PointXYZI (x,y,z,pad,intensity,...) and a packed 16 B xyzi are both accepted as they are;
 *   - host buffers are caller-owned; device memory is owned by the handle; one handle = one CUDA stream;
 *     a handle is not re-entrant, different handles are independent (the reference uses these objects from
 *     three host threads: mapOptmization.cpp:1770-1771);
 *   - return value: 0 = ok, < 0 = error (see b2_status). The reference's own `false` returns (fewer than 50
 *     correspondences, not converged) are out-parameters, not errors;
 *   - there is no CPU fallback: without a CUDA device every compute call returns B2_ERR_CUDA.
 */
#ifndef B2REG_H
#define B2REG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B2_OK = 0,
    B2_ERR_ARG = -1,          /* null pointer / bad size / bad stride */
    B2_ERR_CUDA = -2,         /* CUDA runtime error (message in b2_last_error) */
    B2_ERR_STATE = -3,        /* call order (e.g. iterate before set_map) */
    B2_ERR_CAPACITY = -4,     /* output buffer too small */
    B2_ERR_TOO_LARGE = -5,    /* index grid would not fit the configured budget */
    B2_ERR_NCCL = -6
} b2_status;

/* layout helper: where intensity sits for a given stride (PCL PointXYZI: 16, packed xyzi: 12) */
#define B2_INTENSITY_OFFSET(stride_bytes) ((stride_bytes) >= 32 ? 16 : 12)

int         b2_version(void);
const char* b2_last_error(void);                 /* thread-local, never NULL */
int         b2_device_count(void);
int         b2_set_device(int ordinal);          /* device used by handles created afterwards on this thread */
/* device buffers released by handles are pooled for reuse; this returns the pool to the driver */
int         b2_trim_memory(void);
/* number of CUDA kernels this library has launched in this process so far (bench.py's gpu_launches) */
unsigned long long b2_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * VoxelGrid — replaces pcl::VoxelGrid<PointT>::{setLeafSize,setMinimumPointsNumberPerVoxel,setInputCloud,filter}
 *   liosam_ws/src/LIO-SAM/src/featureExtraction.cpp:233-234
 *   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:719,879,928,932,960,965
 *   Calibration_Tookit/multi_lidar/src/multi_lidar_calibration/src/multi_lidar_calibrator.cpp:113-121
 *   heading_ws/src/src/PointCloudProcessing.cpp:23-30
 * Semantics: PCL 1.10 applyFilter — bbox, idx = sum((floor(p*inv_leaf) - min_b) * mul), one centroid of
 * (x,y,z,intensity) per voxel with >= min_points points, output in ascending voxel index; when
 * dx*dy*dz > INT32_MAX the filter refuses: *refused = 1 and the output is a copy of the input.
 * Within a voxel the float sum runs in ascending input index. */
typedef struct b2_voxel_s* b2_voxel_t;
int b2_voxel_create(b2_voxel_t* out);
int b2_voxel_destroy(b2_voxel_t h);
int b2_voxel_set_leaf_size(b2_voxel_t h, float lx, float ly, float lz);
int b2_voxel_set_min_points_per_voxel(b2_voxel_t h, unsigned min_points);
/* n_fields: 3 (xyz only, PointXYZ) or 4 (xyz + intensity). out may alias nothing; out_capacity in points.
 * voxel_of_point (optional, n ints): linear voxel index of every input point (-1 when refused). */
int b2_voxel_filter(b2_voxel_t h, const void* in, size_t in_stride, size_t n, int n_fields,
                    void* out, size_t out_stride, size_t out_capacity, size_t* n_out,
                    int* refused, int32_t* voxel_of_point);
/* device ms of the last b2_voxel_filter, copies excluded (CUDA events on the handle's stream) */
int b2_voxel_last_gpu_ms(b2_voxel_t h, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Batched nearest-neighbour index — replaces pcl::KdTreeFLANN<PointType>::{setInputCloud,nearestKSearch}
 *   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:1289-1290 (build), :987, :1079 (k = 5 queries)
 * A per-query GPU call is useless, so the query is batched: m queries in, m*k (index, squared distance)
 * out, ascending by (distance, index). Distances are ((dx*dx)+dy*dy)+dz*dz in float without FMA, i.e.
 * FLANN's L2_Simple. The index is a uniform grid with cell edge >= max_dist, so results are exact for every
 * neighbour closer than max_dist — the only ones the reference consumes (it rejects a feature unless
 * sqDis[4] < 1.0, mapOptmization.cpp:993,1089). Slots with no neighbour inside max_dist hold index -1 and
 * +inf. Equal distances are ordered by the smaller index (FLANN's order is traversal dependent). */
typedef struct b2_knn_s* b2_knn_t;
int b2_knn_create(b2_knn_t* out, float max_dist /* 1.0f for LIO-SAM */);
int b2_knn_destroy(b2_knn_t h);
int b2_knn_set_input_cloud(b2_knn_t h, const void* pts, size_t stride, size_t n);
int b2_knn_nearest_k_search(b2_knn_t h, const void* queries, size_t stride, size_t m, int k /* 1..8 */,
                            int32_t* indices, float* sq_dists);

/* ------------------------------------------------------------------------------------------------
 * Scan-to-map Levenberg-Marquardt — replaces, as one fused device op per iteration,
 *   mapOptimization::cornerOptimization        liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:974-1064
 *   mapOptimization::surfOptimization          :1066-1135
 *   mapOptimization::combineOptimizationCoeffs :1137-1156
 *   mapOptimization::LMOptimization            :1158-1280
 * and the loop of scan2MapOptimization :1282-1310 (guards included, transformUpdate excluded).
 * pose = transformTobeMapped = (roll, pitch, yaw, x, y, z), float, in/out. */
typedef struct b2_s2m_s* b2_s2m_t;
typedef struct {
    int   edge_feature_min_valid_num;   /* 10   utility.h:219 */
    int   surf_feature_min_valid_num;   /* 100  utility.h:220 */
    int   max_iterations;               /* 30   mapOptmization.cpp:1292 */
    int   min_correspondences;          /* 50   mapOptmization.cpp:1178 */
    float knn_max_dist;                 /* 1.0  (sqDis[4] < 1.0, :993,:1089) */
    float degenerate_eigen_threshold;   /* 100  :1239 */
    int   max_batch;                    /* scans that can be registered against the map in one call (>= 1) */
} b2_s2m_params;
void b2_s2m_default_params(b2_s2m_params* p);
int  b2_s2m_create(b2_s2m_t* out, const b2_s2m_params* params /* NULL = defaults */);
int  b2_s2m_destroy(b2_s2m_t h);
/* replaces kdtreeCornerFromMap->setInputCloud / kdtreeSurfFromMap->setInputCloud (:1289-1290).
 * b2_s2m_set_map and b2_s2m_set_scan return as soon as the caller's buffers have been read (they may be freed or
 * overwritten at once); the index builds and copies may still be running on the handle's streams. Every later call on the
 * handle is ordered behind them, and the solve / iterate calls end with the one synchronisation of the sequence, so errors
 * of the queued work surface there. */
int  b2_s2m_set_map(b2_s2m_t h, const void* corner, size_t corner_stride, size_t n_corner,
                    const void* surf, size_t surf_stride, size_t n_surf);
/* laserCloudCornerLastDS / laserCloudSurfLastDS of the current scan (:955-967) */
int  b2_s2m_set_scan(b2_s2m_t h, const void* corner, size_t corner_stride, size_t n_corner,
                     const void* surf, size_t surf_stride, size_t n_surf);
/* One loop body: corner+surf+combine+LMOptimization(iter). The six sines/cosines of LMOptimization and the
 * pose matrix are evaluated on the host with the C library (as the reference does) and shipped with the pose.
 * n_sel = laserCloudSelNum; *ran = 0 when n_sel < min_correspondences (LMOptimization returned false early). */
int  b2_s2m_iterate(b2_s2m_t h, float pose[6], int iter, int* n_sel, int* ran, int* converged,
                    int* degenerate, float matP[36]);
/* Whole loop on the device, no host round trip between iterations. Returns B2_OK and *status_not_enough = 1
 * (pose untouched) when the feature-count guards of :1287 fail. pose_history (optional): max_iterations*6. */
int  b2_s2m_solve(b2_s2m_t h, float pose[6], int max_iterations, int* iters_done, int* converged,
                  int* degenerate, float matP[36], int* not_enough_features, float* pose_history);
/* isDegenerate / matP are members that persist across scans in the reference (:136,:234) */
int  b2_s2m_set_state(b2_s2m_t h, int degenerate, const float matP[36]);
/* Introspection of the last iteration (parity tests): which = 0 corner / 1 surf; any pointer may be NULL.
 * knn_idx/knn_d2: n*5, coeff: n*4 (coeffSel entries, valid where flag != 0), flag: n bytes. */
int  b2_s2m_get_pass(b2_s2m_t h, int which, int32_t* knn_idx, float* knn_d2, float* coeff, uint8_t* flag);
int  b2_s2m_get_normal_equations(b2_s2m_t h, float AtA[36], float AtB[6], float X[6]);

/* Batched variant: B independent scans (ragged) against the one resident map, each with its own pose.
 * corner_offsets / surf_offsets: B+1 prefix offsets into the concatenated feature arrays. */
int  b2_s2m_set_scan_batch(b2_s2m_t h, int batch, const void* corner, size_t corner_stride, const int32_t* corner_offsets,
                           const void* surf, size_t surf_stride, const int32_t* surf_offsets);
/* poses: B*6 in/out; iters_done/converged/degenerate: B each (optional). */
int  b2_s2m_solve_batch(b2_s2m_t h, float* poses, int max_iterations, int* iters_done, int* converged,
                        int* degenerate);

/* Timing hooks for bench.py: GPU time in ms of the last solve / solve_batch call, measured with CUDA events
 * on the handle's own stream (torch.cuda.Event only sees torch's stream), and the kernel launches it made. */
int  b2_s2m_last_gpu_ms(b2_s2m_t h, float* ms, int* launches);
/* The two setInputCloud calls of mapOptmization.cpp:1289-1290 once more, on the map clouds the last b2_s2m_set_map /
 * b2_s2m_set_map_from_localmap left in device memory (the reference rebuilds both kd-trees for every scan). Returns with the
 * builds queued, like b2_s2m_set_map. b2_s2m_last_step_gpu_ms: device time from the start of that rebuild to the end of the
 * solve that followed it (CUDA events) = one reference step "index build + LM loop" with the inputs resident in HBM. */
/* Measurement hook: while enabled, batched solves count the map points their neighbour search actually loads (all scans, all
 * iterations of the solve); last_count returns the total of the last counted solve. Used for the roofline numerator. */
int  b2_s2m_count_candidates(b2_s2m_t h, int enable, unsigned long long* last_count);
int  b2_s2m_rebuild_map_index(b2_s2m_t h);
int  b2_s2m_last_step_gpu_ms(b2_s2m_t h, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Local-map assembly — replaces mapOptimization::extractCloud and the containers behind it
 *   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp
 *     :1512-1524  cornerCloudKeyFrames.push_back / surfCloudKeyFrames.push_back      -> b2_localmap_add_keyframe
 *     :899-938    extractCloud(cloudToExtract): transformPointCloud through the laserCloudMapContainer cache, `+=`
 *                 concatenation in visiting order, downSizeFilterCorner / downSizeFilterSurf, cache cleared above
 *                 1000 entries                                                         -> b2_localmap_extract
 *     :1591       laserCloudMapContainer.clear() after correctPoses()                -> b2_localmap_set_pose + b2_localmap_clear_cache
 *     :1289-1290  kdtree*FromMap->setInputCloud(laserCloud*FromMapDS)                -> b2_s2m_set_map_from_localmap
 * Key-frame clouds are uploaded once and stay in HBM; the assembled, downsampled map never crosses PCIe. The caller
 * keeps the key-pose bookkeeping (which key frames are near: extractNearby, :862-897, and the distance test of :905) and
 * passes the indices in the order the reference visits them — the VoxelGrid centroid sums depend on that order.
 * pose6 = (roll, pitch, yaw, x, y, z) of cloudKeyPoses6D. */
typedef struct b2_localmap_s* b2_localmap_t;
int b2_localmap_create(b2_localmap_t* out, float mapping_corner_leaf_size /* 0.2 params.yaml:64 */, float mapping_surf_leaf_size /* 0.4 :65 */);
int b2_localmap_destroy(b2_localmap_t h);
int b2_localmap_add_keyframe(b2_localmap_t h, const void* corner, size_t corner_stride, size_t n_corner,
                             const void* surf, size_t surf_stride, size_t n_surf, const float pose6[6], int* key_index);
int b2_localmap_num_keyframes(b2_localmap_t h, int* n);
int b2_localmap_set_pose(b2_localmap_t h, int key_index, const float pose6[6]);
int b2_localmap_clear_cache(b2_localmap_t h);
int b2_localmap_extract(b2_localmap_t h, const int32_t* key_indices, int n_keys, size_t* n_corner_ds, size_t* n_surf_ds);
/* which: 0 laserCloudCornerFromMap, 1 laserCloudSurfFromMap, 2 laserCloudCornerFromMapDS, 3 laserCloudSurfFromMapDS;
 * out may be NULL to query the size */
int b2_localmap_get(b2_localmap_t h, int which, void* out, size_t stride_bytes, size_t capacity, size_t* n);
int b2_localmap_last_gpu_ms(b2_localmap_t h, float* ms, size_t* n_cached);
int b2_s2m_set_map_from_localmap(b2_s2m_t s2m, b2_localmap_t h);

/* ------------------------------------------------------------------------------------------------
 * Per-scan front end — replaces, for one incoming scan,
 *   ImageProjection::projectPointCloud + deskewPoint + cloudExtraction   liosam_ws/src/LIO-SAM/src/imageProjection.cpp:446-598
 *   FeatureExtraction::calculateSmoothness + markOccludedPoints + extractFeatures   featureExtraction.cpp:81-238
 * Input points are the reference's PointXYZIRT wire struct, 32 B: x@0 y@4 z@8 intensity@16 ring(u16)@20 time(f32)@24
 * (imageProjection.cpp:4-15). The gyro table is the one imuDeskewInfo builds (:305-362): n_imu entries, the last one is
 * imuPointerCur; n_imu <= 1 means imuAvailable == false (no deskew). deskew_flag = -1 disables deskew (:491).
 * Outputs are the cloud_info arrays of msg/cloud_info.msg. Any output pointer may be NULL. */
typedef struct b2_scan_s* b2_scan_t;
typedef struct {
    int   n_scan;                    /* 16    utility.h:197 */
    int   horizon_scan;              /* 1800  utility.h:198 */
    int   downsample_rate;           /* 1     utility.h:199 */
    float lidar_min_range;           /* 1.0   utility.h:200 */
    float lidar_max_range;           /* 1000  utility.h:201 */
    float edge_threshold;            /* 1.0   config/params.yaml:57 (code default 0.1, utility.h:217) */
    float surf_threshold;            /* 0.1   config/params.yaml:58 */
    float odometry_surf_leaf_size;   /* 0.4   config/params.yaml:63 */
} b2_scan_params;
void b2_scan_default_params(b2_scan_params* p);
int  b2_scan_create(b2_scan_t* out, const b2_scan_params* params /* NULL = defaults */);
int  b2_scan_destroy(b2_scan_t h);
/* extracted_xyzi / point_col_ind / point_range: capacity n_scan*horizon_scan entries; start/end_ring_index: n_scan;
 * range_mat: n_scan*horizon_scan floats (FLT_MAX = empty); full_cloud: n_scan*horizon_scan*4 floats (NaN = empty). */
int  b2_scan_project(b2_scan_t h, const void* xyzirt, size_t n,
                     const double* imu_time, const double* imu_rot_x, const double* imu_rot_y, const double* imu_rot_z, int n_imu,
                     double time_scan_cur, int deskew_flag,
                     size_t* n_extracted, float* extracted_xyzi, int32_t* point_col_ind, float* point_range,
                     int32_t* start_ring_index, int32_t* end_ring_index, float* range_mat, float* full_cloud);
/* Runs on the device-resident result of the last b2_scan_project. corner_xyzi: capacity n_scan*120 points
 * (cornerCloud, push order), corner_index: their indices into the extracted cloud; surf_xyzi: capacity
 * n_scan*horizon_scan points (surfaceCloud after the per-ring VoxelGrid); curvature / picked_after_mask / label:
 * n_extracted entries (cloudCurvature, cloudNeighborPicked after markOccludedPoints, final cloudLabel). */
int  b2_scan_extract_features(b2_scan_t h, size_t* n_corner, float* corner_xyzi, int32_t* corner_index,
                              size_t* n_surf, float* surf_xyzi, float* curvature, int32_t* picked_after_mask, int32_t* label);
int  b2_scan_last_gpu_ms(b2_scan_t h, float* ms);
/* Replaces ImageProjection::imuDeskewInfo (imageProjection.cpp:305-362) — host only (<= ~60 dependent
 * double additions per scan). The IMU queue is passed as arrays in arrival order, after imuConverter (:203-204):
 * stamp[n] = header.stamp.toSec(), orientation_xyzw[4n] (may be NULL: rpy_init untouched), angular_velocity[3n].
 * Out: n_popped = messages dropped from the queue front (stamp < time_scan_cur - 0.01, :309-315); the table
 * imu_time / imu_rot_{x,y,z}[0 .. n_table) exactly as b2_scan_project takes it (n_table - 1 == imuPointerCur after :356);
 * imu_available (:361); rpy_init = cloudInfo.imu{Roll,Pitch,Yaw}Init from the newest message at or before time_scan_cur
 * (tf getRPY in double, narrowed to the message's float32; untouched if there is none). capacity = length of the table
 * arrays (the reference's queueLength = 2000, :45, unchecked there): B2_ERR_CAPACITY instead of writing past it. */
int  b2_imu_deskew_info(const double* stamp, const double* orientation_xyzw, const double* angular_velocity, int n,
                        double time_scan_cur, double time_scan_end,
                        double* imu_time, double* imu_rot_x, double* imu_rot_y, double* imu_rot_z, int capacity,
                        int* n_table, int* n_popped, int* imu_available, float rpy_init[3]);

/* ------------------------------------------------------------------------------------------------
 * Rigid transform of a cloud — replaces mapOptimization::transformPointCloud (mapOptmization.cpp:286-305)
 * pose6 = (roll, pitch, yaw, x, y, z); matrix from pcl::getTransformation evaluated on the host in float. */
int  b2_transform_cloud(const void* in, size_t in_stride, size_t n, const float pose6[6],
                        void* out, size_t out_stride);

/* ------------------------------------------------------------------------------------------------
 * Device-resident fp64 cloud — the object behind Multi_LiCa's Open3D calls
 *   Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py
 *     :306-307  o3d.geometry.PointCloud(self.source.pcd)          -> b2_cloud_create + b2_cloud_set_points
 *     :314-315  pcd.voxel_down_sample(voxel_size)                 -> b2_cloud_voxel_down_sample
 *     :327-328  pcd.estimate_normals()  (KDTreeSearchParamKNN(30)) -> b2_cloud_estimate_normals(c, 30)
 *     :347-358  pcd.transform(T)                                  -> b2_cloud_transform
 * Points are N x 3 doubles (np.asarray(pcd.points)); they stay in HBM between the calls.
 * voxel_down_sample: voxel = floor((p - (min_bound - v/2)) / v), one double mean per voxel. Open3D returns the voxels
 * in hash-map order; this library returns them in ascending (z, y, x) voxel order (compare as a set keyed by voxel).
 * estimate_normals: covariance of the knn nearest neighbours (the point included, ties to the smaller index),
 * eigenvector of the smallest eigenvalue; sign: first non-zero of (z, y, x) positive; (0,0,1) with < 3 neighbours. */
typedef struct b2_cloud_s* b2_cloud_t;
int b2_cloud_create(b2_cloud_t* out);
int b2_cloud_destroy(b2_cloud_t c);
int b2_cloud_set_points(b2_cloud_t c, const double* xyz, size_t n);
/* float clouds (pcl::PointXYZ 16 B / PointXYZI 32 B strides) are widened on the device */
int b2_cloud_set_points_f32(b2_cloud_t c, const void* base, size_t stride_bytes, size_t n);
int b2_cloud_size(b2_cloud_t c, size_t* n, int* has_normals);
int b2_cloud_get_points(b2_cloud_t c, double* xyz);
int b2_cloud_get_normals(b2_cloud_t c, double* normals);
int b2_cloud_set_normals(b2_cloud_t c, const double* normals);
/* *out is a new cloud owned by the caller. voxel_rank_of_point (optional, n ints): for every input point the position
 * of its voxel in the output (-1 for non-finite points). */
int b2_cloud_voxel_down_sample(b2_cloud_t c, double voxel_size, b2_cloud_t* out, int32_t* voxel_rank_of_point);
int b2_cloud_estimate_normals(b2_cloud_t c, int knn /* <= 32 */);
int b2_cloud_transform(b2_cloud_t c, const double T[16] /* row-major 4x4 */);
int b2_cloud_last_gpu_ms(b2_cloud_t c, float* ms);

/* ------------------------------------------------------------------------------------------------
 * NCCL communicator for the sharded registrations: one process per GPU; rank 0 creates the id,
 * the host program ships its 128 bytes to the other ranks (torch.distributed / MPI / a file), every rank calls
 * b2_comm_create after b2_set_device. libnccl.so.2 is resolved at run time (B2_NCCL_LIB overrides the name). */
typedef struct b2_comm_s* b2_comm_t;
int b2_comm_unique_id(unsigned char id[128]);
int b2_comm_create(b2_comm_t* out, const unsigned char id[128], int rank, int world);
int b2_comm_destroy(b2_comm_t c);
int b2_comm_rank(b2_comm_t c, int* rank, int* world);
int b2_comm_allreduce_f64(b2_comm_t c, double* values, size_t n);     /* host values, summed over ranks in place */
/* Sharded set-up of a large registration (config C5; every rank holds the same host cloud, as the ranks of the reference's
 * batch scheduler would, multi_lidar_calibrator.py:202-219): upload 1/world of the rows per rank + all-gather over NVLink;
 * normals of 1/world of the points per rank + all-gather (bit-identical to b2_cloud_estimate_normals). */
int b2_cloud_set_points_sharded(b2_cloud_t c, const double* xyz, size_t n, b2_comm_t comm);
int b2_cloud_estimate_normals_sharded(b2_cloud_t c, int knn, b2_comm_t comm);

/* ------------------------------------------------------------------------------------------------
 * Generalized ICP — replaces o3d.pipelines.registration.registration_generalized_icp(source, target, max_corr, init,
 *   TransformationEstimationForGeneralizedICP(epsilon), ICPConvergenceCriteria(rel_fitness, rel_rmse, max_iteration))
 *   Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py:331-340
 * Both clouds need normals (Calibration.py:327-328 always estimates them); the per-point covariance is Open3D's
 * R diag(epsilon,1,1) R^T built from the normal. The whole loop runs on the device: correspondences (exact 1-NN
 * within max_correspondence_distance, ties to the smaller index), J^T J / J^T r, LDL^T solve, T <- [Rz Ry Rx | t] T,
 * stop when |d fitness| < relative_fitness and |d inlier_rmse| < relative_rmse or after max_iteration updates.
 * sums[30] layout: J^T J upper triangle row-major (21), J^T r (6), n_corr, sum of squared point distances, sum r^2. */
typedef struct b2_gicp_s* b2_gicp_t;
typedef struct {
    double max_correspondence_distance;   /* 1.0    Multi_LiCa config/params.yaml:52 */
    double epsilon;                       /* 0.005  :58 */
    double relative_fitness;              /* 1e-7   :59 */
    double relative_rmse;                 /* 1e-7   :60 */
    int    max_iteration;                 /* 100    :61 */
} b2_gicp_params;
void b2_gicp_default_params(b2_gicp_params* p);
int  b2_gicp_create(b2_gicp_t* out, const b2_gicp_params* params /* NULL = defaults */);
int  b2_gicp_destroy(b2_gicp_t h);
int  b2_gicp_set_params(b2_gicp_t h, const b2_gicp_params* params);
int  b2_gicp_set_target(b2_gicp_t h, b2_cloud_t target);    /* builds the target index; the cloud may be destroyed afterwards */
int  b2_gicp_set_source(b2_gicp_t h, b2_cloud_t source);
/* sharded registration: index only rows [begin, end) of the source on this rank (all of them are processed here; the sums are
 * all-reduced over the communicator of b2_gicp_set_shard, the fitness refers to the whole cloud) */
int  b2_gicp_set_source_slice(b2_gicp_t h, b2_cloud_t source, size_t begin, size_t end);
/* the same with spatially dense shards: blocks of 4096 consecutive points of the cloud's Morton order (left by estimate_normals),
 * block b on rank b mod world. The preferred deal for large registrations (better locality than row slices of a shuffled cloud). */
int  b2_gicp_set_source_blocks(b2_gicp_t h, b2_cloud_t source, int rank, int world);
/* Source sharding for one registration spread over `world` GPUs: the (cell-sorted) source is dealt to the ranks in
 * blocks of 4096 consecutive points, round-robin (block b belongs to rank b mod world), so every rank sees the same
 * mix of easy and hard regions; the 30 sums are all-reduced over `comm` every iteration and every rank solves the same
 * 6x6 system, so all ranks return the same T. world = 1 (the default) needs no communicator. With world > 1 and
 * comm = NULL only b2_gicp_linearize works and returns this shard's local sums. */
/* Fused linearise + exchange for sharded registrations (no NCCL call and no epilogue launch per iteration): each rank exports
 * its exchange area (CUDA IPC handle, 64 bytes), the host program gathers the handles of all ranks in rank order and hands them
 * to every rank; the last CTA of k_gicp_linearize then stores its 30 sums into every peer over NVLink (240 B per peer and
 * iteration) and waits for theirs. Needs P2P access between the GPUs; when a peer cannot be mapped b2_gicp_set_peers fails
 * and the NCCL all-reduce of b2_gicp_set_shard stays in use. Call after b2_gicp_set_shard. */
int  b2_gicp_peer_handle(b2_gicp_t h, unsigned char handle[64]);
int  b2_gicp_set_peers(b2_gicp_t h, int rank, int world, const unsigned char* handles /* world x 64 bytes */);
/* both in one: the handles travel over the communicator given to b2_gicp_set_shard (one 64-byte all-gather) */
int  b2_gicp_exchange_setup(b2_gicp_t h);
int  b2_gicp_set_shard(b2_gicp_t h, int rank, int world, b2_comm_t comm);
/* one evaluation at T (parity tests): sums as above; corr (optional, n_source ints) = target index or -1 */
int  b2_gicp_linearize(b2_gicp_t h, const double T[16], double sums[30], int32_t* corr);
int  b2_gicp_align(b2_gicp_t h, const double init[16], double T_out[16], double* fitness, double* inlier_rmse,
                   int* iterations, int* converged);
/* fitness / inlier_rmse of every evaluation of the last align (the first one is the evaluation at init) */
int  b2_gicp_get_history(b2_gicp_t h, double* fitness, double* inlier_rmse, int capacity, int* n_evaluations);
int  b2_gicp_last_gpu_ms(b2_gicp_t h, float* ms, int* launches);
/* device time of every evaluation of the last align (the all-reduce and the update of a sharded run included) */
int  b2_gicp_get_evaluation_ms(b2_gicp_t h, float* ms, int capacity, int* n_evaluations);
int  b2_gicp_index_info(b2_gicp_t h, double* target_cell_edge, double* target_points_per_cell,
                        uint32_t* shard_points, uint32_t* shard_block_points);

/* ------------------------------------------------------------------------------------------------
 * NDT — replaces pcl::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ> as used by
 *   Calibration_Tookit/multi_lidar/src/multi_lidar_calibration/src/multi_lidar_calibrator.cpp:35-72
 *     :37-41  setTransformationEpsilon / setStepSize / setResolution / setMaximumIterations
 *     :43-44  setInputSource / setInputTarget
 *     :62     align(output, guess)
 *     :64-65  hasConverged / getFitnessScore / getTransformationProbability
 *     :69,71  getFinalTransformation
 * Semantics: PCL 1.10 ndt.hpp / voxel_grid_covariance.hpp (defaults: resolution 1.0, step 0.1, epsilon 0.1, 35 iterations,
 * outlier ratio 0.55, >= 6 points per voxel). The calibrator constructs a fresh object on every 10 Hz tick (:35); the
 * PCL-shaped shim keeps one handle per target cloud. Matrices are row-major 4x4 floats (Eigen::Matrix4f is column-major:
 * the shim transposes). Point arrays are (base, stride_bytes, n) with x,y,z first (pcl::PointXYZ: stride 16). */
typedef struct b2_ndt_s* b2_ndt_t;
int b2_ndt_create(b2_ndt_t* out);
int b2_ndt_destroy(b2_ndt_t h);
int b2_ndt_set_transformation_epsilon(b2_ndt_t h, double epsilon);
int b2_ndt_set_step_size(b2_ndt_t h, double step_size);
int b2_ndt_set_resolution(b2_ndt_t h, float resolution);
int b2_ndt_set_maximum_iterations(b2_ndt_t h, int max_iterations);
int b2_ndt_set_input_target(b2_ndt_t h, const void* xyz, size_t stride_bytes, size_t n);
int b2_ndt_set_input_source(b2_ndt_t h, const void* xyz, size_t stride_bytes, size_t n);
/* out_cloud (optional): n_source points of out_stride bytes, the source transformed by the final transformation */
int b2_ndt_align(b2_ndt_t h, const float guess[16], void* out_cloud, size_t out_stride);
int b2_ndt_has_converged(b2_ndt_t h, int* converged);
int b2_ndt_get_final_transformation(b2_ndt_t h, float T[16]);
int b2_ndt_get_fitness_score(b2_ndt_t h, double* score);
int b2_ndt_get_transformation_probability(b2_ndt_t h, double* probability);
int b2_ndt_get_final_num_iteration(b2_ndt_t h, int* iterations);
/* parity hooks: the voxel statistics of the target (ascending voxel index; n_points = -1 marks a rejected covariance)
 * and one derivative pass at the pose vector p = (x, y, z, rx, ry, rz); hessian may be NULL */
int b2_ndt_get_voxels(b2_ndt_t h, size_t capacity, size_t* n_voxels, int32_t* voxel_index, int32_t* n_points, float* centroid_xyz,
                      double* mean, double* inverse_covariance, int32_t min_b[3], int32_t div_b[3]);
int b2_ndt_derivatives(b2_ndt_t h, const double p[6], double* score, double gradient[6], double hessian[36], long long* n_pairs);
/* device time of the last align (or target build), kernel launches and derivative passes it made, (point, voxel) pairs of the last pass */
int b2_ndt_last_gpu_ms(b2_ndt_t h, float* ms, int* launches, int* evaluations, long long* pairs_last);

/* ------------------------------------------------------------------------------------------------
 * Nearest-neighbour registration error + yaw grid search — replaces, in SensorsCalibration's
 * lidar2lidar auto-calibration, Calibration_Tookit/SensorsCalibration/lidar2lidar/auto_calib/src/registration_icp.cpp
 *   :51-52   pcl::KdTreeFLANN kdtree; kdtree.setInputCloud(tgt_ngcloud_)          -> b2_nnerr_set_target
 *   :78-100  CalculateICPError(kdtree, init_guess, yaw): sum of squared 1-NN distances of the transformed source
 *                                                                                   -> b2_nnerr_evaluate
 *   :49-76   RegistrationByICP(init_guess, transform): 37 evaluations on a shrinking yaw grid -> b2_nnerr_yaw_search
 * T is a row-major 4x4 double (Eigen::Matrix4d is column-major: transpose). The reference's GetDeltaT converts its argument
 * from degrees although the caller passes radians; the search reproduces that as written. */
typedef struct b2_nnerr_s* b2_nnerr_t;
int b2_nnerr_create(b2_nnerr_t* out);
int b2_nnerr_destroy(b2_nnerr_t h);
int b2_nnerr_set_target(b2_nnerr_t h, const void* pts, size_t stride_bytes, size_t n);
int b2_nnerr_set_source(b2_nnerr_t h, const void* pts, size_t stride_bytes, size_t n);
int b2_nnerr_evaluate(b2_nnerr_t h, const double T[16], double* dist_sum, size_t* n_found);
int b2_nnerr_yaw_search(b2_nnerr_t h, const double init_guess[16], double T_out[16], double* best_yaw, double* min_error, int* evaluations);
int b2_nnerr_last_gpu_ms(b2_nnerr_t h, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Point-to-point ICP — replaces pcl::IterativeClosestPoint<PointType, PointType> of the loop-closure
 * thread, liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:559-586
 *   :561-565  setMaxCorrespondenceDistance(historyKeyframeSearchRadius*2) / setMaximumIterations(100) /
 *             setTransformationEpsilon(1e-6) / setEuclideanFitnessEpsilon(1e-6) / setRANSACIterations(0)
 *   :568-571  setInputSource / setInputTarget / align            :573  hasConverged, getFitnessScore
 *   :580,:586 getFinalTransformation
 * Semantics: PCL 1.10 (icp.hpp, DefaultConvergenceCriteria, TransformationEstimationSVD); the handle starts with PCL's
 * defaults (10 iterations, no thresholds). Matrices are row-major 4x4 floats. guess / out_cloud may be NULL. */
typedef struct b2_icp_s* b2_icp_t;
int b2_icp_create(b2_icp_t* out);
int b2_icp_destroy(b2_icp_t h);
int b2_icp_set_max_correspondence_distance(b2_icp_t h, double distance);
int b2_icp_set_maximum_iterations(b2_icp_t h, int n);
int b2_icp_set_transformation_epsilon(b2_icp_t h, double epsilon);
int b2_icp_set_euclidean_fitness_epsilon(b2_icp_t h, double epsilon);
int b2_icp_set_ransac_iterations(b2_icp_t h, int n /* only 0 */);
int b2_icp_set_input_source(b2_icp_t h, const void* pts, size_t stride_bytes, size_t n);
int b2_icp_set_input_target(b2_icp_t h, const void* pts, size_t stride_bytes, size_t n);
int b2_icp_align(b2_icp_t h, const float guess[16], void* out_cloud, size_t out_stride);
int b2_icp_has_converged(b2_icp_t h, int* converged);
int b2_icp_get_fitness_score(b2_icp_t h, double* score);
int b2_icp_get_final_transformation(b2_icp_t h, float T[16]);
int b2_icp_get_final_num_iteration(b2_icp_t h, int* iterations);
int b2_icp_last_gpu_ms(b2_icp_t h, float* ms, int* launches);

/* ------------------------------------------------------------------------------------------------
 * cloud_info wire format and device hand-offs between the three LIO-SAM stages
 *   msg/cloud_info.msg:1-35                           the message (ROS1 serialisation, little endian)
 *   utility.h:286-295          publishCloud           pcl::toROSMsg of a pcl::PointXYZI cloud: height 1, width n, fields
 *                                                     x@0 y@4 z@8 intensity@16 (FLOAT32), point_step 32, data = the points
 *   imageProjection.cpp:600-605    publishClouds      stage 0: arrays + cloud_deskewed        -> b2_scan_write_cloud_info(.., 0, ..)
 *   featureExtraction.cpp:66-79    laserCloudInfoHandler (fromROSMsg of cloud_deskewed)       -> b2_scan_set_from_cloud_info
 *   featureExtraction.cpp:240-258  freeCloudInfoMemory + publishFeatureCloud: stage 1: arrays emptied, cloud_deskewed kept,
 *                                                     cloud_corner + cloud_surface added      -> b2_scan_write_cloud_info(.., 1, ..)
 *   mapOptmization.cpp:237-247     laserCloudInfoHandler (fromROSMsg x 2) + :940-958 downsampleCurrentScan
 *                                                                                             -> b2_s2m_set_scan_downsampled
 *   the same hand-off without a message when both stages live in one process                  -> b2_s2m_set_scan_from_front_end
 * Point records are written with data[3] = 1.0f (the PointXYZI constructor) and zero padding after intensity (PCL leaves
 * those 12 bytes uninitialised). The key_frame_* clouds are default-constructed (empty), as the reference publishes them. */
typedef struct {
    uint32_t seq, stamp_sec, stamp_nsec;       /* header of the message (cloudHeader) */
    const char* frame_id;                      /* header.frame_id; NULL = "" */
    const char* cloud_frame_id;                /* lidarFrame, the frame_id publishCloud stamps on the clouds; NULL = "" */
    int64_t imu_available, odom_available;
    float imu_roll_init, imu_pitch_init, imu_yaw_init;
    float initial_guess_x, initial_guess_y, initial_guess_z, initial_guess_roll, initial_guess_pitch, initial_guess_yaw;
} b2_cloud_info_meta;
typedef struct {                               /* one sensor_msgs/PointCloud2 inside the message; pointers into the caller's buffer */
    const void* data;                          /* NOT aligned; usable as the (base, stride = point_step, n = width * height) argument */
    uint32_t height, width, point_step, row_step, n_fields;
    int32_t off_x, off_y, off_z, off_intensity;   /* byte offsets of the FLOAT32 fields, -1 if absent */
    uint8_t is_bigendian, is_dense;
} b2_cloud2_view;
typedef struct {
    uint32_t seq, stamp_sec, stamp_nsec;
    const char* frame_id; uint32_t frame_id_len;        /* not NUL-terminated */
    const void* start_ring_index; uint32_t n_start_ring_index;   /* int32, NOT aligned */
    const void* end_ring_index;   uint32_t n_end_ring_index;
    const void* point_col_ind;    uint32_t n_point_col_ind;
    const void* point_range;      uint32_t n_point_range;         /* float32 */
    int64_t imu_available, odom_available;
    float imu_roll_init, imu_pitch_init, imu_yaw_init;
    float initial_guess_x, initial_guess_y, initial_guess_z, initial_guess_roll, initial_guess_pitch, initial_guess_yaw;
    b2_cloud2_view cloud_deskewed, cloud_corner, cloud_surface, key_frame_cloud, key_frame_color, key_frame_poses, key_frame_map;
} b2_cloud_info_view;
/* Host-only walk over a serialised lio_sam/cloud_info; B2_ERR_ARG on a truncated or malformed message. */
int b2_cloud_info_parse(const void* msg, size_t n_bytes, b2_cloud_info_view* view);
/* Serialise the handle's device-resident result: stage 0 after b2_scan_project, stage 1 after b2_scan_extract_features.
 * One packing kernel and one device-to-host copy into `out`; out may be NULL to query *n_bytes. */
int b2_scan_write_cloud_info(b2_scan_t h, const b2_cloud_info_meta* meta, int stage, void* out, size_t capacity, size_t* n_bytes);
/* The featureExtraction side: load a stage-0 message into the handle (one upload, one unpacking kernel) so that
 * b2_scan_extract_features runs on it. */
int b2_scan_set_from_cloud_info(b2_scan_t h, const void* msg, size_t n_bytes, size_t* n_extracted);
/* laserCloudCornerLast / laserCloudSurfLast -> downSizeFilterCorner / downSizeFilterSurf -> the optimiser's scan, with the
 * downsampled clouds never leaving the device. The two VoxelGrid handles carry the leaf sizes (mappingCornerLeafSize,
 * mappingSurfLeafSize). n_*_ds: laserCloudCornerLastDSNum / laserCloudSurfLastDSNum. */
int b2_s2m_set_scan_downsampled(b2_s2m_t h, b2_voxel_t ds_corner, const void* corner, size_t corner_stride, size_t n_corner,
                                b2_voxel_t ds_surf, const void* surf, size_t surf_stride, size_t n_surf,
                                size_t* n_corner_ds, size_t* n_surf_ds);
int b2_s2m_set_scan_from_front_end(b2_s2m_t h, b2_scan_t scan, b2_voxel_t ds_corner, b2_voxel_t ds_surf,
                                   size_t* n_corner_ds, size_t* n_surf_ds);
/* laserCloudCornerLastDS / laserCloudSurfLastDS of the last set_scan* call (which = 0 corner, 1 surf), packed xyzi. */
int b2_s2m_get_scan(b2_s2m_t h, int which, float* xyzi, size_t capacity, size_t* n);

/* ------------------------------------------------------------------------------------------------
 * Multi-lidar fusion front end — replaces, in PointClouds_Fusion,
 *   fusion_pointclouds/src/fusion_pointcloud/src/fusion_pointclouds.cpp
 *     :62-73   pcl::transformPointCloud(*pc_local_k, *pc_trans_k, T_k.matrix())  (double matrix)  -> b2_fusion_add_cloud(.., T_k)
 *     :80-89   pc_fusion_local = pc_local_1 [+ pc_trans_4] [+ pc_trans_3] + pc_trans_2            -> the order of the add_cloud calls
 *     :93-108  passthroughFiter(external bounds) / conditionFiter(internal bounds)               -> b2_fusion_set_*_bounds
 *   lidar_fusion/src/src/lidar_fusion.cpp:239-252, :333-334 (transform one cloud, cloud1 + cloud2)
 * Points are PCL records (stride >= 16: pcl::PointXYZI 32 B, or packed xyzi 16 B). External bounds keep min <= v <= max
 * (float limits, non-finite points dropped, as pcl::PassThrough); internal bounds keep what lies outside the box
 * (pcl::ConditionOr of GT / LT comparisons). Output order = input order. */
typedef struct b2_fusion_s* b2_fusion_t;
int b2_fusion_create(b2_fusion_t* out);
int b2_fusion_destroy(b2_fusion_t h);
int b2_fusion_clear(b2_fusion_t h);                                  /* forget the added clouds */
int b2_fusion_add_cloud(b2_fusion_t h, const void* pts, size_t stride_bytes, size_t n, const double T[16] /* NULL: untransformed */);
int b2_fusion_set_external_bounds(b2_fusion_t h, int enabled, const double min_xyz[3], const double max_xyz[3]);
int b2_fusion_set_internal_bounds(b2_fusion_t h, int enabled, const double min_xyz[3], const double max_xyz[3]);
int b2_fusion_run(b2_fusion_t h, size_t* n_fused, size_t* n_out);
int b2_fusion_get(b2_fusion_t h, void* out, size_t stride_bytes, size_t capacity, size_t* n);   /* out may be NULL to query n */
int b2_fusion_device_cloud(b2_fusion_t h, const void** d_xyzi, size_t* n);                      /* packed float4 in device memory */
int b2_fusion_last_gpu_ms(b2_fusion_t h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* B2REG_H */
