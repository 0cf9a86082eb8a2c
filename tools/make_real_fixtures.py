"""Writes tests/golden/real_lidar2lidar_0001.npz from the real clouds the reference ships (TEST INFRASTRUCTURE).

Source files (identical bytes in both places, see the sha256 recorded in the fixture):
  Calibration_Tookit/Multi_LiCa/data/demo/lidar_{1,2,3}.pcd                       (Multi_LiCa demo, config/demo.yaml)
  Calibration_Tookit/SensorsCalibration/lidar2lidar/auto_calib/data/0001/{top,left,right}.pcd
and the start the reference itself uses for them:
  Calibration_Tookit/SensorsCalibration/lidar2lidar/auto_calib/data/0001/initial_extrinsic.txt:1-6
The published accuracy of the Multi_LiCa pipeline (its only quantitative statement about this path) is the table
  Calibration_Tookit/Multi_LiCa/evaluation/config.yaml:4-13  (ground truth vs calibration, x y z [m] roll pitch yaw [deg]);
its per-axis differences are stored as `published_abs_error` and their maxima as `envelope` (0.0424 m, 0.345 deg).

/root/reference does not exist on the GPU box, so the clouds are committed as float32 arrays (the PCD files store float32)
together with this script. lidar_1 keeps its `ring` and per-point `time` (seconds from the first return of the sweep), which
is what LIO-SAM's imageProjection consumes (PointXYZIRT, imageProjection.cpp:4-15).

    python tools/make_real_fixtures.py
"""
import hashlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pcd_io  # noqa: E402

REF = "/root/reference/Calibration_Tookit"
DEMO = REF + "/Multi_LiCa/data/demo"
CASE = REF + "/SensorsCalibration/lidar2lidar/auto_calib/data/0001"
OUT = os.path.join(HERE, "..", "tests", "golden", "real_lidar2lidar_0001.npz")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    out = {}
    for name, twin in (("lidar_1", "top"), ("lidar_2", "left"), ("lidar_3", "right")):
        a, b = f"{DEMO}/{name}.pcd", f"{CASE}/{twin}.pcd"
        assert sha(a) == sha(b), (a, b)
        d = pcd_io.read_pcd(a)
        xyz = np.stack([d["x"], d["y"], d["z"]], 1).astype(np.float32)
        assert np.isfinite(xyz).all()
        out[name] = xyz
        out[name + "_sha256"] = np.array(sha(a))
        if name == "lidar_1":
            out["lidar_1_intensity"] = d["intensity"].astype(np.float32)
            out["lidar_1_ring"] = d["ring"].astype(np.uint8)
            t = d["timestamp"]
            out["lidar_1_time"] = (t - t.min()).astype(np.float32)
    # initial_extrinsic.txt: "(Roll,Pitch,Yaw,tx,ty,tz): r p y tx ty tz" per device, degrees / metres
    rows = []
    for line in open(f"{CASE}/initial_extrinsic.txt"):
        m = re.match(r"\(Roll,Pitch,Yaw,tx,ty,tz\):\s*(.*)", line.strip())
        if m:
            rows.append([float(v) for v in m.group(1).split()])
    out["initial_extrinsic_rpy_deg_xyz"] = np.array(rows)          # device 0 (top), 1 (left), 2 (right)
    # evaluation/config.yaml:4-13
    gt = np.array([[-0.05003868754137674, -0.6182657287315831, -0.2210218828997021, 19.761386585960206, 9.671599526442137, 4.528509610943743],
                   [0.018654221998691264, -1.2213014405123956, -0.44237105646436303, 39.9427218516336, -0.48189320093417537, 1.220711041394239],
                   [-2.692777651895696, -0.6935691501242932, -0.11539450889299188, 19.95269089910927, -11.158756732625704, -3.0853232386037823]])
    cal = np.array([[-0.04623345928427628, -0.6083580787632855, -0.2261904420354155, 19.76077565640768, 9.6105862444547, 4.48426248738788],
                    [-0.015582894322383505, -1.2184022475297067, -0.43754838940491786, 39.96339416907645, -0.3281780704703811, 1.2074037966430986],
                    [-2.6503996565781494, -0.6947256501155331, -0.1061083588396785, 19.61996156314735, -10.813415564912178, -3.052343294964041]])
    txt = open(REF + "/Multi_LiCa/evaluation/config.yaml").read()
    for v in np.concatenate([gt.ravel(), cal.ravel()]):
        assert repr(float(v)) in txt or str(v) in txt, v           # the numbers above are the file's
    err = np.abs(cal - gt)
    out["published_abs_error"] = err
    out["envelope"] = np.array([err[:, :3].max(), err[:, 3:].max()])   # [m], [deg]
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes; envelope", out["envelope"])


if __name__ == "__main__":
    main()
