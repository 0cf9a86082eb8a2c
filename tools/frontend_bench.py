#!/usr/bin/env python
"""Per-scan front-end workload of BASELINE.json (config C2) — used by bench.py and runnable alone.

A synthetic 128-ring x 1024-column scan (131 072 returns before drop-out) of the city-block scene with the gyro table of
C2 (omega(t) = (0.3 sin 7t, 0.2 cos 5t, 0.8) rad/s at 500 Hz); one step = ImageProjection::projectPointCloud + deskew +
cloudExtraction, then FeatureExtraction::calculateSmoothness + markOccludedPoints + extractFeatures (per-ring VoxelGrid
included), host buffers in and out. Unit: input points per second (Mpts/s).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T0 = 1000.0


def omega(t):
    return (0.3 * np.sin(7 * t), 0.2 * np.cos(5 * t), 0.8)


def c2_inputs():
    from multi_sensor_slam_tookit_b200 import synth
    scene = synth.CityBlock()
    raw = synth.ring_scan(scene, (0.0, 0.0, 0.3, 0.0, -24.0, 1.8), n_rings=128, n_cols=1024, elev_deg=(-22.5, 22.5),
                          seed=synth.MASTER_SEED + 2, omega=omega)
    return raw, synth.imu_table(T0, 0.1, omega)


def run_c2(steps=50, warmup=5, inputs=None):
    from multi_sensor_slam_tookit_b200.frontend import ScanFrontEnd
    raw, imu = inputs or c2_inputs()
    fe = ScanFrontEnd(128, 1024)
    for _ in range(warmup):
        fe.projectPointCloud(raw, imu=imu, timeScanCur=T0); fe.extractFeatures()
    dev_p = dev_f = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        p = fe.projectPointCloud(raw, imu=imu, timeScanCur=T0); dev_p += fe.lastGpuMs()
        f = fe.extractFeatures(); dev_f += fe.lastGpuMs()
    wall = time.perf_counter() - t0
    n_in, m = len(raw), len(p["extracted"])
    return dict(points_in=int(n_in), extracted=int(m), corner=int(len(f["corner"])), surf=int(len(f["surf"])), steps=steps,
                project_gpu_ms=dev_p / steps, features_gpu_ms=dev_f / steps, e2e_ms=1e3 * wall / steps,
                mpts_per_s=n_in * steps / ((dev_p + dev_f) * 1e-3) / 1e6, e2e_mpts_per_s=n_in * steps / wall / 1e6,
                # 52 B per input point (projection + deskew) + 20 B per extracted point (curvature + masks)
                algorithmic_bytes=52.0 * n_in + 20.0 * m)


def cpu_c2(inputs, budget_s=6.0):
    from oracle import pyoracle as O
    raw, imu = inputs
    reps, t0 = 0, time.perf_counter()
    tp = tf = 0.0
    while True:
        a = time.perf_counter(); o = O.project(raw, 128, 1024, imu=imu, t_cur=T0); b = time.perf_counter()
        O.extract_features(o, 1.0, 0.1, 0.4, stable=True); c = time.perf_counter()
        tp += b - a; tf += c - b; reps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return dict(reps=reps, project_ms=1e3 * tp / reps, features_ms=1e3 * tf / reps, mpts_per_s=len(raw) * reps / (tp + tf) / 1e6)


if __name__ == "__main__":
    inp = c2_inputs()
    out = {"c2": run_c2(inputs=inp)}
    if "--cpu" in sys.argv:
        out["cpu_c2"] = cpu_c2(inp)
    print(json.dumps(out), flush=True)
