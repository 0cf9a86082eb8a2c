"""include/b2reg_pcl_shim.hpp is what a maintainer of the reference drops into the PCL call sites. PCL and Eigen are not in
this image, so the header is syntax-checked against minimal stand-ins (tests/stubs/pcl): every shim class is instantiated,
which type-checks each call into the C ABI of include/b2reg.h (argument counts, pointer and stride types)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TU = r'''
#include "b2reg_pcl_shim.hpp"
int use() {
    using P = pcl::PointXYZI;
    pcl::PointCloud<P>::Ptr a(new pcl::PointCloud<P>()), b(new pcl::PointCloud<P>());
    pcl::PointCloud<P> out;
    b2shim::VoxelGrid<P> vg; vg.setLeafSize(0.4f, 0.4f, 0.4f); vg.setMinimumPointsNumberPerVoxel(2); vg.setInputCloud(a); vg.filter(out);
    b2shim::VoxelGrid<pcl::PointXYZ> vg3; (void)vg3;
    b2shim::ScanToMap s2m; s2m.setInputMap(*a, *b); s2m.setInputScan(*a, *b);
    float pose[6] = {0, 0, 0, 0, 0, 0}; bool deg = false; bool conv = s2m.optimize(pose, deg);
    b2shim::NormalDistributionsTransform<P, P> ndt;
    ndt.setTransformationEpsilon(0.01); ndt.setStepSize(0.1); ndt.setResolution(1.0f); ndt.setMaximumIterations(400);
    ndt.setInputSource(a); ndt.setInputTarget(b); ndt.align(out, Eigen::Matrix4f::Identity());
    Eigen::Matrix4f T = ndt.getFinalTransformation();
    b2shim::IterativeClosestPoint<P, P> icp;
    icp.setMaxCorrespondenceDistance(30.0); icp.setMaximumIterations(100); icp.setTransformationEpsilon(1e-6);
    icp.setEuclideanFitnessEpsilon(1e-6); icp.setRANSACIterations(0); icp.setInputSource(a); icp.setInputTarget(b); icp.align(out);
    T = icp.getFinalTransformation();
    return (conv ? 1 : 0) + (ndt.hasConverged() ? 1 : 0) + (icp.hasConverged() ? 1 : 0) + (int)ndt.getFitnessScore() + (int)icp.getFitnessScore() +
           ndt.getFinalNumIteration() + (int)ndt.getTransformationProbability() + (int)T(0, 0);
}
'''


def test_pcl_shim_header_type_checks_against_the_c_abi():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "tu.cpp")
        with open(src, "w") as f:
            f.write(TU)
        r = subprocess.run([cxx, "-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                            "-I", os.path.join(ROOT, "tests", "stubs"), src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
