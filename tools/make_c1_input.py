"""Builds tests/golden/c1_input.npz — the config-1 workload of BASELINE.json (SURVEY.md §8d "C1").

A synthetic VLP-16 scan (16 x 1800) and a 100 000-point local feature map, both produced by running the
CPU ORACLE's front end (projection -> curvature -> feature selection -> VoxelGrid) on ray-cast scans of the
city-block scene, exactly as LIO-SAM would assemble them:
  key-frame clouds = per-scan corner/surf features after downsampleCurrentScan (mapOptmization.cpp:955-967),
  local map        = key frames within 50 m, transformed, concatenated, VoxelGrid 0.2 / 0.4 (mapOptmization.cpp:862-953),
  then the nearest 100 000 points to the current pose are kept (natural corner/surf split recorded).
This script is test infrastructure (it imports oracle/); bench.py and the tests only read the .npz it writes.
Run: python tools/make_c1_input.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_sensor_slam_tookit_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

SEED = synth.MASTER_SEED + 1
EDGE_TH, SURF_TH = 1.0, 0.1                 # config/params.yaml:57-58
ODOM_SURF_LEAF, MAP_CORNER_LEAF, MAP_SURF_LEAF = 0.4, 0.2, 0.4   # params.yaml:63-65
KEYFRAME_RADIUS = 50.0                      # params.yaml:79
MAP_POINTS = 100_000


def scan_features(scene, pose, seed):
    raw = synth.ring_scan(scene, pose, seed=seed)
    proj = O.project(raw, 16, 1800, imu=None)
    feat = O.extract_features(proj, EDGE_TH, SURF_TH, ODOM_SURF_LEAF, stable=True)
    corner_ds = O.voxel_grid(feat["corner"], MAP_CORNER_LEAF)["out"]
    surf_ds = O.voxel_grid(feat["surf"], MAP_SURF_LEAF)["out"]
    return raw, corner_ds, surf_ds


def build(spacing=1.0):
    scene = synth.CityBlock(synth.MASTER_SEED)
    poses = synth.street_loop(2, 1, spacing)
    cur = poses[0].copy()
    near = np.linalg.norm(poses[:, 3:6] - cur[3:6], axis=1) <= KEYFRAME_RADIUS
    poses = poses[near]
    mc, ms = [], []
    for k, p in enumerate(poses):
        _, c, s = scan_features(scene, p, SEED * 1000 + k)
        pf = p.astype(np.float32)
        mc.append(O.transform_cloud(c, pf))
        ms.append(O.transform_cloud(s, pf))
    mapc = O.voxel_grid(np.concatenate(mc), MAP_CORNER_LEAF)["out"]
    maps = O.voxel_grid(np.concatenate(ms), MAP_SURF_LEAF)["out"]
    return scene, cur, len(poses), mapc, maps


def main():
    spacing = 1.0
    while True:
        scene, cur, nkf, mapc, maps = build(spacing)
        print(f"spacing {spacing}: {nkf} key frames -> corner {len(mapc)} surf {len(maps)}")
        if len(mapc) + len(maps) >= MAP_POINTS:
            break
        spacing /= 2
    # nearest MAP_POINTS to the current pose, order within each cloud preserved (ascending voxel index)
    d = np.concatenate([np.linalg.norm(mapc[:, :3] - cur[3:6], axis=1), np.linalg.norm(maps[:, :3] - cur[3:6], axis=1)])
    thr = np.partition(d, MAP_POINTS - 1)[MAP_POINTS - 1]
    keep = d <= thr
    extra = int(keep.sum()) - MAP_POINTS
    if extra > 0:
        keep[np.flatnonzero(d == thr)[:extra]] = False
    mapc, maps = mapc[keep[:len(mapc)]], maps[keep[len(mapc):]]
    # the scan to register: taken 0.5 m further along the street, so it is not itself a key frame
    truth = cur.copy()
    truth[3] += 0.5
    raw, sc, ss = scan_features(scene, truth, SEED * 1000 + 999_999)
    # initial guess = truth perturbed by (0.10, -0.08, 0.03) m and (0.5, -0.4, 1.0) deg  
    guess = truth.copy()
    guess[3:6] += np.array([0.10, -0.08, 0.03])
    guess[0:3] += np.deg2rad([0.5, -0.4, 1.0])
    out = os.path.join(ROOT, "tests", "golden", "c1_input.npz")
    np.savez_compressed(out, scan_corner=sc.astype(np.float32), scan_surf=ss.astype(np.float32),
                        map_corner=mapc.astype(np.float32), map_surf=maps.astype(np.float32),
                        pose_truth=truth.astype(np.float32), pose_guess=guess.astype(np.float32),
                        keyframe_spacing=np.float32(spacing), n_keyframes=np.int32(nkf))
    print(f"wrote {out}: scan corner {len(sc)} surf {len(ss)}; map corner {len(mapc)} surf {len(maps)} "
          f"({100.0 * len(mapc) / MAP_POINTS:.1f}% corner); raw scan {len(raw)} returns; {os.path.getsize(out) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
