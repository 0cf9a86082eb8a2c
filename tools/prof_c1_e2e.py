"""Wall-clock breakdown of the C1 e2e step (developer tool)."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_input.npz"))
d = {k: z[k] for k in z.files}
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
mc, ms, sc, ss = pin(d["map_corner"]), pin(d["map_surf"]), pin(d["scan_corner"]), pin(d["scan_surf"])
guess = d["pose_guess"].copy()
g = ScanToMapOptimizer()
T = {"set_map": 0.0, "set_scan": 0.0, "solve": 0.0, "total": 0.0}
N = 400
best = 1e9
for i in range(N + 50):
    t0 = time.perf_counter(); g.setInputMap(mc, ms); t1 = time.perf_counter(); g.setInputScan(sc, ss); t2 = time.perf_counter()
    g.transformTobeMapped = guess.copy(); r = g.scan2MapOptimization(30, want_matP=False); t3 = time.perf_counter()
    if i >= 50:
        T["set_map"] += t1 - t0; T["set_scan"] += t2 - t1; T["solve"] += t3 - t2; T["total"] += t3 - t0; best = min(best, t3 - t0)
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print({k: round(v / N * 1e6, 1) for k, v in T.items()}, "us per step; best", round(best * 1e6, 1), "iters", r["iters"], "gpu_ms", round(g.lastGpuMs()[0] * 1e3, 1), "|", clk)
