// Device-resident fp64 point cloud — the object the Open3D calls of Multi_LiCa operate on
// (Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py:306-340: PointCloud copy,
// voxel_down_sample, estimate_normals, registration_generalized_icp) and the PointXYZ clouds of the NDT calibrator.
// Points stay in HBM between those calls; nothing goes back to the host unless the caller asks for it.
#pragma once
#include "b2_common.cuh"
#include "b2_comm.cuh"
#include "b2_gridd.cuh"
#include "b2_bvh.cuh"

struct b2_cloud_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    size_t n = 0;
    b2::DevBuf xyz;            // double[3n], caller's order (Open3D points_)
    b2::DevBuf nrm;            // double[3n], caller's order (Open3D normals_)
    bool has_normals = false;
    b2::DevBuf work;           // scratch for downsampling
    b2::PinBuf pin;
    float last_ms = 0.f;
    // estimate_normals leaves its search kernel queued on `stream` (everything that consumes the normals is ordered behind it on
    // the same stream or synchronises it): the two clouds of a registration pair are two single-wave launches that overlap
    // instead of running back to back. The BVH and the event pair therefore outlive the call.
    b2::BvhIndex bvh;
    b2::DevBuf gathered;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ms_pending = false;
};

namespace b2 {
// min/max of n device points (3 doubles each) -> host mn[3], mx[3]; returns B2_OK; empty/non-finite clouds give mn > mx
int bbox_f64(const double* d_xyz, size_t n, DevBuf& scratch, cudaStream_t s, double mn[3], double mx[3]);
// cumulants -> covariance -> eigenvector of the smallest eigenvalue, as Open3D's estimate_normals does per point
int estimate_normals_knn(b2_cloud_s* c, int knn, ::b2_comm_s* comm);
}  // namespace b2
