"""Wall-clock breakdown of one C4 pair through the public API (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gicp_bench as B
from multi_sensor_slam_tookit_b200 import gicp, capi
import torch
torch.cuda.init()
clouds = B.c4_clouds()
P = B.C4_PARAMS
def sync():
    torch.cuda.synchronize()
for rep in range(3):
    T = {}
    def tick(name, t0):
        sync(); T[name] = T.get(name, 0) + (time.perf_counter() - t0) * 1e3
    s, t = 1, 0
    init, truth = B.c4_init(s, t)
    t0 = time.perf_counter(); a = gicp.PointCloud(clouds[s]); b = gicp.PointCloud(clouds[t]); tick("upload", t0)
    t0 = time.perf_counter(); sp = a.voxel_down_sample(0.05); tp = b.voxel_down_sample(0.05); tick("voxel_down_sample", t0)
    t0 = time.perf_counter(); sp.estimate_normals(); tick("normals_src", t0); T["normals_src_gpu"] = sp.lastGpuMs()
    t0 = time.perf_counter(); tp.estimate_normals(); tick("normals_tgt", t0); T["normals_tgt_gpu"] = tp.lastGpuMs()
    g = gicp.GeneralizedICP(1.0, 0.005, 1e-7, 1e-7, 100)
    t0 = time.perf_counter(); g.setInputTarget(tp); tick("set_target", t0)
    t0 = time.perf_counter(); g.setInputSource(sp); tick("set_source", t0)
    t0 = time.perf_counter(); r = g.align(init, want_correspondences=False); tick("align", t0)
    T["align_gpu"] = r.gpu_ms; T["iters"] = r.iterations; T["info"] = g.indexInfo()
    t0 = time.perf_counter(); r = g.align(init, want_correspondences=True); tick("align+corr", t0)
    print(rep, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in T.items()}, flush=True)
