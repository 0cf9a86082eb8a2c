// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of pcl::VoxelGrid<PointXYZI>::filter as the reference calls it:
//   liosam_ws/src/LIO-SAM/src/featureExtraction.cpp:233-234 (per-ring surf DS)
//   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:719,879,928,932,960,965 (map / scan DS)
//   Calibration_Tookit/multi_lidar/src/multi_lidar_calibration/src/multi_lidar_calibrator.cpp:113-121
//   heading_ws/src/src/PointCloudProcessing.cpp:23-30 (setMinimumPointsNumberPerVoxel(2))
// PCL is not vendored; this follows the published PCL 1.10 algorithm:
//   inv = 1/leaf (float); bbox; refuse when dx*dy*dz > INT32_MAX (output = input);
//   idx = sum((floor(p*inv) - min_b) * mul); sort by idx; one centroid of x,y,z,intensity per
//   voxel holding >= min_points_per_voxel points; output ascending idx.
// Within-voxel summation order: PCL's std::sort is unstable, so the order is unspecified there.
// The oracle pins it to ascending input index ("stable mode").
// parity unpinned by reference tests (the reference has none); pinned by construction + properties.
#include <vector>
#include <cstdint>
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <limits>

extern "C" {

// in: n points, xyzi interleaved (16 B stride). out: capacity n points.
// voxel_of_point (optional, n ints): linear voxel index per input point (-1 when refused).
// returns number of output points; *refused = 1 when the index would overflow (output = input copy).
int o_voxel_grid(const float* in, int n, float lx, float ly, float lz, unsigned min_points_per_voxel,
                 float* out, int* voxel_of_point, int* refused, int* out_voxel_idx) {
    if (refused) *refused = 0;
    if (n <= 0) return 0;
    const float inv[3] = {1.0f / lx, 1.0f / ly, 1.0f / lz};
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            float v = in[(size_t)i * 4 + d];
            mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
        }
    int64_t dxyz[3];
    for (int d = 0; d < 3; d++) dxyz[d] = (int64_t)((mx[d] - mn[d]) * inv[d]) + 1;
    if (dxyz[0] * dxyz[1] * dxyz[2] > (int64_t)std::numeric_limits<int32_t>::max()) {
        if (refused) *refused = 1;
        for (size_t i = 0; i < (size_t)n * 4; i++) out[i] = in[i];
        if (voxel_of_point) for (int i = 0; i < n; i++) voxel_of_point[i] = -1;
        return n;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int d = 0; d < 3; d++) {
        min_b[d] = (int)std::floor(mn[d] * inv[d]);
        max_b[d] = (int)std::floor(mx[d] * inv[d]);
        div_b[d] = max_b[d] - min_b[d] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    std::vector<std::pair<unsigned, int>> iv(n);
    for (int i = 0; i < n; i++) {
        const float* p = &in[(size_t)i * 4];
        int ijk0 = (int)(std::floor(p[0] * inv[0]) - (float)min_b[0]);
        int ijk1 = (int)(std::floor(p[1] * inv[1]) - (float)min_b[1]);
        int ijk2 = (int)(std::floor(p[2] * inv[2]) - (float)min_b[2]);
        int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
        iv[i] = {(unsigned)idx, i};
        if (voxel_of_point) voxel_of_point[i] = idx;
    }
    std::sort(iv.begin(), iv.end());   // (idx, input index): the stable order
    int total = 0;
    size_t index = 0;
    while (index < iv.size()) {
        size_t i = index + 1;
        while (i < iv.size() && iv[i].first == iv[index].first) ++i;
        if (i - index >= min_points_per_voxel) {
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            for (size_t li = index; li < i; li++) {
                const float* p = &in[(size_t)iv[li].second * 4];
                sx += p[0]; sy += p[1]; sz += p[2]; si += p[3];
            }
            float cnt = (float)(i - index);
            out[(size_t)total * 4 + 0] = sx / cnt;
            out[(size_t)total * 4 + 1] = sy / cnt;
            out[(size_t)total * 4 + 2] = sz / cnt;
            out[(size_t)total * 4 + 3] = si / cnt;
            if (out_voxel_idx) out_voxel_idx[total] = (int)iv[index].first;
            ++total;
        }
        index = i;
    }
    return total;
}

}  // extern "C"
