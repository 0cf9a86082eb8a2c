// b2reg_pcl_shim.hpp — header-only C++ shims with PCL-shaped interfaces on top of the C ABI (b2reg.h).
//
// Drop-in for the call sites of SURVEY.md §8b: the reference keeps its pcl::PointCloud<PointT> containers and calls the
// same member names; the arithmetic runs in libb2reg.so on the GPU. The header needs only <pcl/point_cloud.h> and
// <pcl/point_types.h> from PCL (containers and point structs, no algorithms). It is NOT compiled in this repository's
// tests (PCL is not installed in the build image); INTEGRATION.md shows where each class goes.
#pragma once
#include <stdexcept>
#include <string>
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

}  // namespace b2shim
