// b2reg_pcl_shim.hpp — header-only C++ shims with PCL-shaped interfaces on top of the C ABI (b2reg.h).
//
// Drop-in for the call sites of SURVEY.md §8b: the reference keeps its pcl::PointCloud<PointT> containers and calls the
// same member names; the arithmetic runs in libb2reg.so on the GPU. The header needs only <pcl/point_cloud.h> and
// <pcl/point_types.h> from PCL (containers and point structs, no algorithms). PCL is not installed in the build image:
// tests/test_shim_compiles.py type-checks it against minimal stand-ins (tests/stubs/pcl) and, on the GPU box, compiles and RUNS
// a program through it on the C1 workload (tests/shim_c1_main.cpp); INTEGRATION.md shows where each class goes.
#pragma once
#include <stdexcept>
#include <string>
#include <type_traits>
#include <cstdint>
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

// replaces pcl::KdTreeFLANN<PointT> where the reference issues one nearestKSearch per point inside a loop
// (mapOptmization.cpp:987,1079; the loop bodies of :978-1063 and :1070-1134 when only the search is moved to the GPU).
// A per-query GPU call is useless, so the batched overload takes all queries of the loop at once; the single-point overload
// keeps PCL's signature (returns the number of neighbours found) for the few call sites outside a loop. Exact k-NN within
// `max_dist` (the reference consumes nothing beyond 1 m, :993,1089); neighbours farther than that are reported as -1 / inf.
// radiusSearch on key poses (:871,460,624 — a few thousand points) stays on pcl::KdTreeFLANN.
template <typename PointT>
class KdTreeFLANN {
public:
    explicit KdTreeFLANN(float max_dist = 1.0f) { check(b2_knn_create(&h_, max_dist), "b2_knn_create"); }
    ~KdTreeFLANN() { b2_knn_destroy(h_); }
    KdTreeFLANN(const KdTreeFLANN&) = delete;
    void setInputCloud(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) {
        check(b2_knn_set_input_cloud(h_, cloud->points.data(), sizeof(PointT), cloud->size()), "setInputCloud");
    }
    // all queries of one loop: k_indices / k_sqr_distances get queries.size() * k entries, row per query, ascending distance
    void nearestKSearch(const pcl::PointCloud<PointT>& queries, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
        k_indices.resize(queries.size() * static_cast<size_t>(k));
        k_sqr_distances.resize(queries.size() * static_cast<size_t>(k));
        static_assert(sizeof(int) == sizeof(int32_t), "int32 indices");
        check(b2_knn_nearest_k_search(h_, queries.points.data(), sizeof(PointT), queries.size(), k,
                                      reinterpret_cast<int32_t*>(k_indices.data()), k_sqr_distances.data()), "nearestKSearch");
    }
    int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
        k_indices.resize(k); k_sqr_distances.resize(k);
        check(b2_knn_nearest_k_search(h_, &point, sizeof(PointT), 1, k, reinterpret_cast<int32_t*>(k_indices.data()), k_sqr_distances.data()), "nearestKSearch");
        int found = 0;
        while (found < k && k_indices[found] >= 0) found++;
        k_indices.resize(found); k_sqr_distances.resize(found);
        return found;
    }
private:
    b2_knn_t h_ = nullptr;
};

// replaces the per-scan bodies of ImageProjection (imageProjection.cpp:180-195: imuDeskewInfo inside deskewInfo(),
// projectPointCloud(), cloudExtraction()) and FeatureExtraction (featureExtraction.cpp:66-79: calculateSmoothness(),
// markOccludedPoints(), extractFeatures()). RawPoint is the node's PointXYZIRT (32 bytes, imageProjection.cpp:4-15); the
// outputs are the arrays of lio_sam/cloud_info plus the three clouds the nodes publish.
struct CloudInfoArrays {                              // msg/cloud_info.msg:4-8
    std::vector<int32_t> startRingIndex, endRingIndex, pointColInd;
    std::vector<float> pointRange;
};
struct ImuQueueView {                                 // the node's std::deque<sensor_msgs::Imu>, after imuConverter, as arrays
    const double* stamp = nullptr; const double* orientation_xyzw = nullptr; const double* angular_velocity = nullptr; int n = 0;
};
class ScanFrontEnd {
public:
    explicit ScanFrontEnd(const b2_scan_params* p = nullptr) {
        if (p) prm_ = *p; else b2_scan_default_params(&prm_);
        check(b2_scan_create(&h_, &prm_), "b2_scan_create");
        imuTime_.resize(2000); imuRotX_.resize(2000); imuRotY_.resize(2000); imuRotZ_.resize(2000);      // queueLength, imageProjection.cpp:45
    }
    ~ScanFrontEnd() { b2_scan_destroy(h_); }
    ScanFrontEnd(const ScanFrontEnd&) = delete;
    // imuDeskewInfo (:305-362): returns cloudInfo.imuAvailable; n_popped messages are to be popped from the node's queue
    bool imuDeskewInfo(const ImuQueueView& q, double timeScanCur, double timeScanEnd, int& n_popped, float rpyInit[3]) {
        int avail = 0;
        check(b2_imu_deskew_info(q.stamp, q.orientation_xyzw, q.angular_velocity, q.n, timeScanCur, timeScanEnd, imuTime_.data(), imuRotX_.data(),
                                 imuRotY_.data(), imuRotZ_.data(), static_cast<int>(imuTime_.size()), &nImu_, &n_popped, &avail, rpyInit), "imuDeskewInfo");
        return avail != 0;
    }
    // projectPointCloud + cloudExtraction (:521-598): laserCloudIn -> extractedCloud + cloud_info arrays. deskewFlag as :491.
    template <typename RawPoint>
    void projectPointCloud(const pcl::PointCloud<RawPoint>& laserCloudIn, double timeScanCur, int deskewFlag,
                           pcl::PointCloud<pcl::PointXYZI>& extractedCloud, CloudInfoArrays& info) {
        static_assert(sizeof(RawPoint) == 32, "PointXYZIRT is 32 bytes: x y z _ intensity ring time _");
        const size_t cells = static_cast<size_t>(prm_.n_scan) * prm_.horizon_scan;
        xyzi_.resize(cells * 4);
        info.pointColInd.resize(cells); info.pointRange.resize(cells);
        info.startRingIndex.resize(prm_.n_scan); info.endRingIndex.resize(prm_.n_scan);
        size_t m = 0;
        check(b2_scan_project(h_, laserCloudIn.points.data(), laserCloudIn.size(), imuTime_.data(), imuRotX_.data(), imuRotY_.data(), imuRotZ_.data(),
                              nImu_, timeScanCur, deskewFlag, &m, xyzi_.data(), info.pointColInd.data(), info.pointRange.data(),
                              info.startRingIndex.data(), info.endRingIndex.data(), nullptr, nullptr), "projectPointCloud");
        info.pointColInd.resize(m); info.pointRange.resize(m);
        unpack(xyzi_.data(), m, extractedCloud);
    }
    // calculateSmoothness + markOccludedPoints + extractFeatures (:81-238) on the projection the handle holds
    void extractFeatures(pcl::PointCloud<pcl::PointXYZI>& cornerCloud, pcl::PointCloud<pcl::PointXYZI>& surfaceCloud) {
        const size_t cells = static_cast<size_t>(prm_.n_scan) * prm_.horizon_scan;
        std::vector<float> c(static_cast<size_t>(prm_.n_scan) * 120 * 4), s(cells * 4);
        std::vector<int32_t> ci(static_cast<size_t>(prm_.n_scan) * 120);
        size_t nc = 0, ns = 0;
        check(b2_scan_extract_features(h_, &nc, c.data(), ci.data(), &ns, s.data(), nullptr, nullptr, nullptr), "extractFeatures");
        unpack(c.data(), nc, cornerCloud); unpack(s.data(), ns, surfaceCloud);
    }
    b2_scan_t handle() const { return h_; }           // for b2_s2m_set_scan_from_front_end when the three stages share a process
private:
    static void unpack(const float* xyzi, size_t n, pcl::PointCloud<pcl::PointXYZI>& out) {
        out.points.resize(n); out.width = static_cast<uint32_t>(n); out.height = 1; out.is_dense = true;
        for (size_t i = 0; i < n; i++) { auto& p = out.points[i]; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; }
    }
    b2_scan_params prm_{};
    b2_scan_t h_ = nullptr;
    std::vector<double> imuTime_, imuRotX_, imuRotY_, imuRotZ_;
    int nImu_ = 0;
    std::vector<float> xyzi_;
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
