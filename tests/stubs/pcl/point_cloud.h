// Minimal stand-in for <pcl/point_cloud.h> (see point_types.h in this directory).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
};
}  // namespace pcl
