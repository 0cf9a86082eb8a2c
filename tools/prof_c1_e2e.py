"""Wall-clock breakdown of the C1 e2e step (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_input.npz"))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
mc, ms, sc, ss = pin(d["map_corner"]), pin(d["map_surf"]), pin(d["scan_corner"]), pin(d["scan_surf"])
g = ScanToMapOptimizer()
T = {"set_map": 0.0, "set_scan": 0.0, "solve": 0.0, "solve_gpu": 0.0}
N = 300
for i in range(N + 20):
    t0 = time.perf_counter(); g.setInputMap(mc, ms); t1 = time.perf_counter(); g.setInputScan(sc, ss); t2 = time.perf_counter()
    g.transformTobeMapped = d["pose_guess"].copy(); r = g.scan2MapOptimization(30, want_matP=False); t3 = time.perf_counter()
    if i >= 20:
        T["set_map"] += t1 - t0; T["set_scan"] += t2 - t1; T["solve"] += t3 - t2; T["solve_gpu"] += g.lastGpuMs()[0] * 1e-3
print({k: round(v / N * 1e6, 1) for k, v in T.items()}, "us per step; iters", r["iters"])
