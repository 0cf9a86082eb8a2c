import sys, os; sys.path.insert(0, '/root/repo')
os.environ["B2_S2M_PROF"] = "1"
import numpy as np
from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer
d = np.load('/root/repo/tests/golden/c1_input.npz')
g = ScanToMapOptimizer(); g.setInputMap(d['map_corner'], d['map_surf']); g.setInputScan(d['scan_corner'], d['scan_surf'])
for rep in range(3):
    g.transformTobeMapped = d['pose_guess'].copy()
    for it in range(3):
        g.LMIteration(it)
