"""Seeded inputs shared by the GICP tests (CPU oracle pins and GPU parity)."""
import numpy as np

from multi_sensor_slam_tookit_b200 import synth


def rig_pose(i):
    """Five lidar mounting poses in the spirit of Calibration_Tookit/multi_lidar/.../cfg/child_topic_list (roll, pitch, yaw, x, y, z)."""
    rig = [(0.0, 0.0, 0.0, 0.0, 0.0, 1.8),
           (0.0, 0.02, 1.5, 1.0, 0.0, 1.8),
           (0.01, 0.0, -1.5, 1.0, -0.6, 1.8),
           (0.0, -0.02, 3.0, -0.8, 0.3, 1.9),
           (-0.01, 0.01, 0.6, 0.4, 0.5, 1.7)]
    return np.array(rig[i])


def pose_matrix(p):
    T = np.eye(4)
    T[:3, :3] = synth.rot_zyx(p[0], p[1], p[2])
    T[:3, 3] = p[3:6]
    return T


def lidar_cloud(i, n_rings=32, n_cols=512, seed=7):
    """Scan of lidar i of the rig, in its own sensor frame, float64 (n, 3)."""
    scene = synth.CityBlock()
    raw = synth.ring_scan(scene, rig_pose(i), n_rings=n_rings, n_cols=n_cols, elev_deg=(-22.5, 22.5),
                          seed=synth.MASTER_SEED + seed + i, noise=0.01, dropout=0.02)
    return np.stack([raw["x"], raw["y"], raw["z"]], 1).astype(np.float64)


def pair_truth(i_src, i_tgt):
    """Transform taking lidar i_src's frame to lidar i_tgt's frame."""
    return np.linalg.inv(pose_matrix(rig_pose(i_tgt))) @ pose_matrix(rig_pose(i_src))


def perturbed(T, d_xyz=(0.10, -0.06, 0.04), d_rpy_deg=(0.8, -0.5, 1.2)):
    r = np.deg2rad(d_rpy_deg)
    D = np.eye(4)
    D[:3, :3] = synth.rot_zyx(r[0], r[1], r[2])
    D[:3, 3] = d_xyz
    return D @ T


def rot_angle(R):
    return float(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)))
