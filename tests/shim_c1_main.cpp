// Runs the reference-side shim classes of include/b2reg_pcl_shim.hpp end to end (test program, built by
// tests/test_shim_compiles.py against the stand-in PCL containers of tests/stubs): the call sequence a maintainer writes in
// mapOptmization.cpp (:955-967 downsampleCurrentScan, :1289-1305 the scan2MapOptimization loop), featureExtraction.cpp (:66-79)
// and imageProjection.cpp (:180-195), on clouds read from raw float32 files. Results go to stdout as text; the Python test
// compares them with the oracle.
//   shim_c1_main <dir>     dir holds map_corner.f32 map_surf.f32 scan_corner.f32 scan_surf.f32 (n x 4 floats) pose.f32 (6 floats) raw.bin (n x 32 B)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "b2reg_pcl_shim.hpp"

struct PointXYZIRT { float x, y, z, pad0, intensity; uint16_t ring; uint16_t pad1; float time; float pad2; };   // imageProjection.cpp:4-15
static_assert(sizeof(PointXYZIRT) == 32, "PointXYZIRT");

static std::vector<float> read_f32(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v(n / 4);
    if (fread(v.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}
static pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_from(const std::vector<float>& v) {
    pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
    c->resize(v.size() / 4);
    for (size_t i = 0; i < c->size(); i++) { auto& p = c->points[i]; p.x = v[4 * i]; p.y = v[4 * i + 1]; p.z = v[4 * i + 2]; p.intensity = v[4 * i + 3]; }
    return c;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string d = argv[1];
    try {
        auto mapCorner = cloud_from(read_f32(d + "/map_corner.f32")), mapSurf = cloud_from(read_f32(d + "/map_surf.f32"));
        auto scanCorner = cloud_from(read_f32(d + "/scan_corner.f32")), scanSurf = cloud_from(read_f32(d + "/scan_surf.f32"));
        std::vector<float> pose = read_f32(d + "/pose.f32");
        // ---- mapOptmization.cpp:955-967 + :1289-1305
        b2shim::VoxelGrid<pcl::PointXYZI> downSizeFilterSurf;
        downSizeFilterSurf.setLeafSize(0.4f, 0.4f, 0.4f);
        pcl::PointCloud<pcl::PointXYZI>::Ptr surfDS(new pcl::PointCloud<pcl::PointXYZI>());
        downSizeFilterSurf.setInputCloud(scanSurf);
        downSizeFilterSurf.filter(*surfDS);
        printf("voxel %zu %zu\n", scanSurf->size(), surfDS->size());
        b2shim::ScanToMap s2m;
        s2m.setInputMap(*mapCorner, *mapSurf);
        s2m.setInputScan(*scanCorner, *scanSurf);
        float transformTobeMapped[6];
        for (int i = 0; i < 6; i++) transformTobeMapped[i] = pose[i];
        bool isDegenerate = false;
        const bool converged = s2m.optimize(transformTobeMapped, isDegenerate);
        printf("pose %d %d %.9g %.9g %.9g %.9g %.9g %.9g\n", converged ? 1 : 0, isDegenerate ? 1 : 0, transformTobeMapped[0], transformTobeMapped[1],
               transformTobeMapped[2], transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5]);
        // ---- the kd-tree alone (loop bodies of :978-1063 with only the search moved): first 64 surf features, k = 5
        b2shim::KdTreeFLANN<pcl::PointXYZI> kdtreeSurfFromMap;
        kdtreeSurfFromMap.setInputCloud(mapSurf);
        pcl::PointCloud<pcl::PointXYZI> q;
        q.resize(64);
        for (int i = 0; i < 64; i++) q.points[i] = mapSurf->points[(size_t)i * 997 % mapSurf->size()];
        std::vector<int> idx; std::vector<float> d2;
        kdtreeSurfFromMap.nearestKSearch(q, 5, idx, d2);
        printf("knn");
        for (size_t i = 0; i < idx.size(); i++) printf(" %d", idx[i]);
        printf("\n");
        std::vector<int> i1; std::vector<float> d1;
        const int found = kdtreeSurfFromMap.nearestKSearch(q.points[3], 5, i1, d1);
        printf("knn1 %d %d %.9g\n", found, i1.empty() ? -1 : i1[0], d1.empty() ? -1.f : d1[0]);
        // ---- imageProjection.cpp:180-195 + featureExtraction.cpp:66-79 on the raw sweep
        FILE* f = fopen((d + "/raw.bin").c_str(), "rb");
        if (f) {
            fseek(f, 0, SEEK_END); long nb = ftell(f); fseek(f, 0, SEEK_SET);
            pcl::PointCloud<PointXYZIRT> laserCloudIn;
            laserCloudIn.resize(nb / 32);
            if (fread(laserCloudIn.points.data(), 1, nb, f) != (size_t)nb) return 2;
            fclose(f);
            b2shim::ScanFrontEnd fe;
            std::vector<double> stamp, gyro;
            for (int k = 0; k < 80; k++) { stamp.push_back(999.98 + k * 0.002); gyro.push_back(0.1); gyro.push_back(-0.05); gyro.push_back(0.4); }
            b2shim::ImuQueueView iq; iq.stamp = stamp.data(); iq.angular_velocity = gyro.data(); iq.n = (int)stamp.size();
            int popped = 0; float rpy[3] = {0, 0, 0};
            const bool imuAvailable = fe.imuDeskewInfo(iq, 1000.0, 1000.1, popped, rpy);
            pcl::PointCloud<pcl::PointXYZI> extractedCloud, cornerCloud, surfaceCloud;
            b2shim::CloudInfoArrays info;
            fe.projectPointCloud(laserCloudIn, 1000.0, 1, extractedCloud, info);
            fe.extractFeatures(cornerCloud, surfaceCloud);
            printf("frontend %d %d %zu %zu %zu %d %d\n", imuAvailable ? 1 : 0, popped, extractedCloud.size(), cornerCloud.size(), surfaceCloud.size(),
                   info.startRingIndex[0], info.endRingIndex[15]);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "shim error: %s\n", e.what());
        return 1;
    }
    return 0;
}
