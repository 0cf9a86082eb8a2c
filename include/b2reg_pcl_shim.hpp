// b2reg_pcl_shim.hpp — header-only C++ shims with PCL-shaped interfaces on top of the C ABI (b2reg.h).
//
// Drop-in for the call sites of SURVEY.md §8b: the reference keeps its pcl::PointCloud<PointT> containers and calls the
// same member names; the arithmetic runs in libb2reg.so on the GPU. The header needs only <pcl/point_cloud.h> and
// <pcl/point_types.h> from PCL (containers and point structs, no algorithms). PCL is not installed in the build image:
// tests/test_shim_compiles.py syntax-checks it against minimal stand-ins (tests/stubs/pcl); INTEGRATION.md shows where
// each class goes.
#pragma once
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "b2reg.h"

namespace b2shim {

inline void check(int rc, const char* what) {
    if (rc != B2_OK) throw std::runtime_error(std::string(what) + ": " + b2_last_error());
}

// replaces pcl::VoxelGrid<PointT> (featureExtraction.cpp:233-234, mapOptmization.cpp:955-967, multi_lidar_calibrator.cpp:113-121)
template <typename PointT>
class VoxelGrid {
public:
    VoxelGrid() { check(b2_voxel_create(&h_), "b2_voxel_create"); }
    ~VoxelGrid() { b2_voxel_destroy(h_); }
    VoxelGrid(const VoxelGrid&) = delete;
    void setLeafSize(float lx, float ly, float lz) { check(b2_voxel_set_leaf_size(h_, lx, ly, lz), "setLeafSize"); }
    void setMinimumPointsNumberPerVoxel(unsigned n) { check(b2_voxel_set_min_points_per_voxel(h_, n), "setMinimumPointsNumberPerVoxel"); }
    void setInputCloud(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) { in_ = cloud; }
    void filter(pcl::PointCloud<PointT>& out) {
        out.clear();
        out.height = 1; out.is_dense = true;
        if (!in_ || in_->empty()) { out.width = 0; return; }                 // PCL: warn + empty output
        const size_t n = in_->size();
        out.points.resize(n);
        size_t m = 0; int refused = 0;
        const int fields = std::is_same<PointT, pcl::PointXYZ>::value ? 3 : 4;
        check(b2_voxel_filter(h_, in_->points.data(), sizeof(PointT), n, fields, out.points.data(), sizeof(PointT), n, &m, &refused, nullptr), "filter");
        out.points.resize(m);
        out.width = static_cast<uint32_t>(m);
    }
private:
    b2_voxel_t h_ = nullptr;
    typename pcl::PointCloud<PointT>::ConstPtr in_;
};

// replaces the two pcl::KdTreeFLANN members + the four optimisation members of mapOptimization (mapOptmization.cpp:974-1310)
class ScanToMap {
public:
    ScanToMap() { check(b2_s2m_create(&h_, nullptr), "b2_s2m_create"); }
    ~ScanToMap() { b2_s2m_destroy(h_); }
    ScanToMap(const ScanToMap&) = delete;
    // kdtreeCornerFromMap->setInputCloud(laserCloudCornerFromMapDS); kdtreeSurfFromMap->setInputCloud(laserCloudSurfFromMapDS);
    void setInputMap(const pcl::PointCloud<pcl::PointXYZI>& corner, const pcl::PointCloud<pcl::PointXYZI>& surf) {
        check(b2_s2m_set_map(h_, corner.points.data(), sizeof(pcl::PointXYZI), corner.size(),
                             surf.points.data(), sizeof(pcl::PointXYZI), surf.size()), "b2_s2m_set_map");
    }
    void setInputScan(const pcl::PointCloud<pcl::PointXYZI>& cornerDS, const pcl::PointCloud<pcl::PointXYZI>& surfDS) {
        check(b2_s2m_set_scan(h_, cornerDS.points.data(), sizeof(pcl::PointXYZI), cornerDS.size(),
                              surfDS.points.data(), sizeof(pcl::PointXYZI), surfDS.size()), "b2_s2m_set_scan");
    }
    // the loop of scan2MapOptimization: returns true when LMOptimization converged; transformTobeMapped updated in place
    bool optimize(float transformTobeMapped[6], bool& isDegenerate, int maxIter = 30) {
        int iters = 0, conv = 0, deg = 0, not_enough = 0;
        check(b2_s2m_solve(h_, transformTobeMapped, maxIter, &iters, &conv, &deg, nullptr, &not_enough, nullptr), "b2_s2m_solve");
        if (!not_enough) isDegenerate = deg != 0;
        return conv != 0;
    }
private:
    b2_s2m_t h_ = nullptr;
};

// replaces pcl::NormalDistributionsTransform<PointSource, PointTarget> (multi_lidar_calibrator.cpp:35-72)
template <typename PointSource, typename PointTarget>
class NormalDistributionsTransform {
public:
    NormalDistributionsTransform() { check(b2_ndt_create(&h_), "b2_ndt_create"); }
    ~NormalDistributionsTransform() { b2_ndt_destroy(h_); }
    NormalDistributionsTransform(const NormalDistributionsTransform&) = delete;
    void setTransformationEpsilon(double e) { check(b2_ndt_set_transformation_epsilon(h_, e), "setTransformationEpsilon"); }
    void setStepSize(double s) { check(b2_ndt_set_step_size(h_, s), "setStepSize"); }
    void setResolution(float r) { check(b2_ndt_set_resolution(h_, r), "setResolution"); }
    void setMaximumIterations(int n) { check(b2_ndt_set_maximum_iterations(h_, n), "setMaximumIterations"); }
    void setInputSource(const typename pcl::PointCloud<PointSource>::ConstPtr& c) {
        src_ = c;
        check(b2_ndt_set_input_source(h_, c->points.data(), sizeof(PointSource), c->size()), "setInputSource");
    }
    void setInputTarget(const typename pcl::PointCloud<PointTarget>::ConstPtr& c) {
        check(b2_ndt_set_input_target(h_, c->points.data(), sizeof(PointTarget), c->size()), "setInputTarget");
    }
    // Eigen::Matrix4f is column-major, the C ABI takes row-major 4x4 floats
    void align(pcl::PointCloud<PointSource>& output, const Eigen::Matrix4f& guess = Eigen::Matrix4f::Identity()) {
        const Eigen::Matrix<float, 4, 4, Eigen::RowMajor> g = guess;
        output.points.resize(src_ ? src_->size() : 0);
        output.width = static_cast<uint32_t>(output.points.size()); output.height = 1; output.is_dense = true;
        check(b2_ndt_align(h_, g.data(), output.points.data(), sizeof(PointSource)), "align");
    }
    bool hasConverged() const { int c = 0; check(b2_ndt_has_converged(h_, &c), "hasConverged"); return c != 0; }
    double getFitnessScore() const { double s = 0; check(b2_ndt_get_fitness_score(h_, &s), "getFitnessScore"); return s; }
    double getTransformationProbability() const { double p = 0; check(b2_ndt_get_transformation_probability(h_, &p), "getTransformationProbability"); return p; }
    int getFinalNumIteration() const { int n = 0; check(b2_ndt_get_final_num_iteration(h_, &n), "getFinalNumIteration"); return n; }
    Eigen::Matrix4f getFinalTransformation() const {
        Eigen::Matrix<float, 4, 4, Eigen::RowMajor> t;
        check(b2_ndt_get_final_transformation(h_, t.data()), "getFinalTransformation");
        return t;
    }
private:
    b2_ndt_t h_ = nullptr;
    typename pcl::PointCloud<PointSource>::ConstPtr src_;
};

// replaces pcl::IterativeClosestPoint<PointSource, PointTarget> as the loop-closure thread configures it (mapOptmization.cpp:559-586)
template <typename PointSource, typename PointTarget>
class IterativeClosestPoint {
public:
    IterativeClosestPoint() { check(b2_icp_create(&h_), "b2_icp_create"); }
    ~IterativeClosestPoint() { b2_icp_destroy(h_); }
    IterativeClosestPoint(const IterativeClosestPoint&) = delete;
    void setMaxCorrespondenceDistance(double d) { check(b2_icp_set_max_correspondence_distance(h_, d), "setMaxCorrespondenceDistance"); }
    void setMaximumIterations(int n) { check(b2_icp_set_maximum_iterations(h_, n), "setMaximumIterations"); }
    void setTransformationEpsilon(double e) { check(b2_icp_set_transformation_epsilon(h_, e), "setTransformationEpsilon"); }
    void setEuclideanFitnessEpsilon(double e) { check(b2_icp_set_euclidean_fitness_epsilon(h_, e), "setEuclideanFitnessEpsilon"); }
    void setRANSACIterations(int n) { check(b2_icp_set_ransac_iterations(h_, n), "setRANSACIterations"); }
    void setInputSource(const typename pcl::PointCloud<PointSource>::ConstPtr& c) {
        src_ = c;
        check(b2_icp_set_input_source(h_, c->points.data(), sizeof(PointSource), c->size()), "setInputSource");
    }
    void setInputTarget(const typename pcl::PointCloud<PointTarget>::ConstPtr& c) {
        check(b2_icp_set_input_target(h_, c->points.data(), sizeof(PointTarget), c->size()), "setInputTarget");
    }
    void align(pcl::PointCloud<PointSource>& output, const Eigen::Matrix4f& guess = Eigen::Matrix4f::Identity()) {
        const Eigen::Matrix<float, 4, 4, Eigen::RowMajor> g = guess;
        output.points.resize(src_ ? src_->size() : 0);
        output.width = static_cast<uint32_t>(output.points.size()); output.height = 1; output.is_dense = true;
        check(b2_icp_align(h_, g.data(), output.points.data(), sizeof(PointSource)), "align");
    }
    bool hasConverged() const { int c = 0; check(b2_icp_has_converged(h_, &c), "hasConverged"); return c != 0; }
    double getFitnessScore() const { double s = 0; check(b2_icp_get_fitness_score(h_, &s), "getFitnessScore"); return s; }
    Eigen::Matrix4f getFinalTransformation() const {
        Eigen::Matrix<float, 4, 4, Eigen::RowMajor> t;
        check(b2_icp_get_final_transformation(h_, t.data()), "getFinalTransformation");
        return t;
    }
private:
    b2_icp_t h_ = nullptr;
    typename pcl::PointCloud<PointSource>::ConstPtr src_;
};

}  // namespace b2shim
