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
    b2shim::KdTreeFLANN<P> kd(1.0f); kd.setInputCloud(a);
    std::vector<int> idx; std::vector<float> d2; kd.nearestKSearch(*a, 5, idx, d2); (void)kd.nearestKSearch(a->points[0], 5, idx, d2);
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


def _build_shim_program(out_dir):
    from multi_sensor_slam_tookit_b200 import capi
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = os.path.join(out_dir, "shim_c1_main")
    lib_dir = os.path.dirname(capi.LIB_PATH)
    r = subprocess.run([cxx, "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "stubs"),
                        os.path.join(ROOT, "tests", "shim_c1_main.cpp"), "-o", exe, "-L", lib_dir, "-l:" + os.path.basename(capi.LIB_PATH),
                        "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_shim_program_compiles_and_links():
    """The reference-side program (every shim class, the call order of the three LIO-SAM nodes) builds and links here, GPU or not."""
    from multi_sensor_slam_tookit_b200 import build
    build.build_lib()
    with tempfile.TemporaryDirectory() as d:
        assert os.path.exists(_build_shim_program(d))


import pytest  # noqa: E402


@pytest.mark.gpu
def test_shim_program_runs_c1_and_matches_oracle(oracle, c1):
    """Compiles AND RUNS the C++ program through b2reg_pcl_shim.hpp on the C1 input: VoxelGrid, ScanToMap, KdTreeFLANN and
    ScanFrontEnd give what the oracle gives (pose within 1e-5 m / 1e-6 rad, voxel count, 5-NN indices, feature counts)."""
    import numpy as np
    from multi_sensor_slam_tookit_b200 import synth
    with tempfile.TemporaryDirectory() as d:
        exe = _build_shim_program(d)
        for k in ("map_corner", "map_surf", "scan_corner", "scan_surf"):
            np.ascontiguousarray(c1[k], np.float32).tofile(os.path.join(d, k + ".f32"))
        np.ascontiguousarray(c1["pose_guess"], np.float32).tofile(os.path.join(d, "pose.f32"))
        scene = synth.CityBlock(synth.MASTER_SEED)
        raw = synth.ring_scan(scene, c1["pose_truth"].astype(np.float64), seed=synth.MASTER_SEED + 4242)
        raw.tofile(os.path.join(d, "raw.bin"))
        r = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.strip().splitlines()}
    # VoxelGrid
    ref_v = oracle.voxel_grid(c1["scan_surf"], 0.4)["out"]
    assert int(out["voxel"][0]) == len(c1["scan_surf"]) and int(out["voxel"][1]) == len(ref_v)
    # scan-to-map
    o = oracle.Scan2Map(4)
    o.set_map(c1["map_corner"], c1["map_surf"]); o.set_scan(c1["scan_corner"], c1["scan_surf"])
    ref = o.solve(c1["pose_guess"])
    pose = np.array([float(v) for v in out["pose"][2:]])
    assert int(out["pose"][0]) == int(ref["converged"])
    assert np.all(np.abs(pose[3:] - ref["pose_hist"][-1][3:]) <= 1e-5) and np.all(np.abs(pose[:3] - ref["pose_hist"][-1][:3]) <= 1e-6)
    # kd-tree (batched and single-point overloads)
    m = np.ascontiguousarray(c1["map_surf"], np.float32)
    q = m[(np.arange(64) * 997) % len(m)]
    oi, od = oracle.knn(m, q, 5)
    assert np.array_equal(np.array([int(v) for v in out["knn"]]).reshape(64, 5), oi)
    assert int(out["knn1"][0]) == 5 and int(out["knn1"][1]) == oi[3, 0] and float(out["knn1"][2]) == od[3, 0]
    # front end: imuDeskewInfo -> projectPointCloud -> extractFeatures
    stamp = 999.98 + np.arange(80) * 0.002
    gyro = np.tile([0.1, -0.05, 0.4], (80, 1))
    info = oracle.imu_deskew_info(stamp, None, gyro, 1000.0, 1000.1)
    proj = oracle.project(raw, 16, 1800, imu=info["imu"], t_cur=1000.0)
    feat = oracle.extract_features(proj)
    fe = out["frontend"]
    assert int(fe[0]) == int(info["imuAvailable"]) and int(fe[1]) == info["n_popped"]
    assert int(fe[2]) == len(proj["extracted"]) and int(fe[3]) == len(feat["corner"]) and int(fe[4]) == len(feat["surf"])
    assert int(fe[5]) == proj["startRingIndex"][0] and int(fe[6]) == proj["endRingIndex"][15]
