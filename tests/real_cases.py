"""The real clouds the reference ships (tests/golden/real_lidar2lidar_0001.npz, written by tools/make_real_fixtures.py from
Multi_LiCa/data/demo/lidar_{1,2,3}.pcd = SensorsCalibration/lidar2lidar/auto_calib/data/0001/{top,left,right}.pcd) and the
parameters the reference runs on them. Shared by the CPU pins of the oracle and the GPU parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Multi_LiCa/config/params.yaml:52-63
GICP = dict(voxel_size=0.05, max_corresp_dist=1.0, epsilon=0.005, rel_fitness=1e-7, rel_rmse=1e-7, max_iterations=100)
# multi_lidar_calibrator.cpp:157-170 (code defaults; BASELINE config 3 pins resolution to 1 m) and :113-121 (child VoxelGrid)
NDT = dict(voxel_size=0.1, resolution=1.0, step_size=0.1, epsilon=0.01, max_iterations=400)

_CACHE = {}


def fixture():
    if "fx" not in _CACHE:
        d = np.load(os.path.join(ROOT, "tests", "golden", "real_lidar2lidar_0001.npz"))
        _CACHE["fx"] = {k: d[k] for k in d.files}
    return _CACHE["fx"]


def rpy_deg_xyz_to_matrix(row):
    """initial_extrinsic.txt row '(Roll,Pitch,Yaw,tx,ty,tz)' [deg, m] -> 4x4, R = Rz(yaw) Ry(pitch) Rx(roll)
    (SensorsCalibration/lidar2lidar/auto_calib/src/run_lidar2lidar.cpp:55-70 degrees -> radians, calibration.cpp:36-49 'rotation = Rz * Ry * Rx')."""
    r, p, y = np.deg2rad(row[:3])
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = row[3:6]
    return T


def initial_guess(name):
    fx = fixture()
    return rpy_deg_xyz_to_matrix(fx["initial_extrinsic_rpy_deg_xyz"][{"lidar_2": 1, "lidar_3": 2}[name]])


def euler_deg(T):
    R = T[:3, :3]
    return np.rad2deg([np.arctan2(R[2, 1], R[2, 2]), np.arcsin(-R[2, 0]), np.arctan2(R[1, 0], R[0, 0])])


def per_axis_difference(Ta, Tb):
    """|x y z| [m] and |roll pitch yaw| [deg] of Ta^-1 Tb — the quantities evaluation/config.yaml tabulates."""
    D = np.linalg.inv(Ta) @ Tb
    return np.abs(D[:3, 3]), np.abs(euler_deg(D))


def perturb(T, xyz, rpy_deg):
    return rpy_deg_xyz_to_matrix(list(rpy_deg) + list(xyz)) @ T


def xyzi(p):
    return np.concatenate([p, np.zeros((len(p), 1), p.dtype)], 1).astype(np.float32)


def lidar1_xyzirt():
    """lidar_1 as LIO-SAM's PointXYZIRT records (imageProjection.cpp:4-15): 64 rings x 0.2 deg, one 0.1 s sweep."""
    from multi_sensor_slam_tookit_b200 import synth
    fx = fixture()
    out = np.zeros(len(fx["lidar_1"]), synth.XYZIRT)
    out["x"], out["y"], out["z"] = fx["lidar_1"].T
    out["intensity"] = fx["lidar_1_intensity"]
    out["ring"] = fx["lidar_1_ring"]
    out["time"] = fx["lidar_1_time"]
    return out
