"""Seeded inputs shared by the NDT tests (config C3 of SURVEY.md §8d, scaled for test time)."""
import numpy as np

import gicp_cases as G


def pair(n_rings=64, n_cols=512, voxel=0.1, oracle=None):
    """Parent (target) scan of lidar 0, child (source) scan of lidar 1 downsampled with pcl::VoxelGrid(voxel) as
    multi_lidar_calibrator.cpp:113-121 does, truth transform child -> parent."""
    tgt = G.lidar_cloud(0, n_rings=n_rings, n_cols=n_cols).astype(np.float32)
    src = G.lidar_cloud(1, n_rings=n_rings, n_cols=n_cols).astype(np.float32)
    if oracle is not None:
        src = oracle.voxel_grid(np.c_[src, np.zeros(len(src), np.float32)], voxel)["out"][:, :3].copy()
    return tgt, src, G.pair_truth(1, 0)


def pose_vector(T):
    """(x, y, z, rx, ry, rz) with R = Rx Ry Rz (the NDT parametrisation)."""
    R = T[:3, :3]
    ry = np.arcsin(np.clip(R[0, 2], -1, 1))
    rx = np.arctan2(-R[1, 2], R[2, 2])
    rz = np.arctan2(-R[0, 1], R[0, 0])
    return np.array([T[0, 3], T[1, 3], T[2, 3], rx, ry, rz])


def guess_from_file_row(x, y, z, yaw, pitch, roll):
    """multi_lidar_calibrator.cpp:50-58: Translation * Rz(yaw) * Ry(pitch) * Rx(roll), float."""
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = (x, y, z)
    return T.astype(np.float32)
