#!/usr/bin/env python
"""GICP workloads of BASELINE.json (configs C4 and C5) — used by bench.py and runnable alone:

    python tools/gicp_bench.py [--c5-points N] [--skip-c4] [--skip-c5] [--cpu]
    torchrun --nproc-per-node 2 ... tools/gicp_bench.py            (sharded C5 / round-robin C4)

C4  five synthetic 64x1024 lidars of one rig, every ordered pair calibrated as Multi_LiCa's fitness mode does
    (multi_lidar_calibrator.py:302-321): voxel_down_sample(0.05) -> estimate_normals() -> registration_generalized_icp
    (max_corr 1.0, epsilon 0.005, 1e-7 / 1e-7, 100 iterations; config/params.yaml:52-63). Pair i runs on rank i % world.
C5  map-to-map GICP of an N-point (default 50 M) surface sample of the city-block scene tiled 4x4 against a noisy,
    rigidly displaced copy; the target is replicated, the source sharded over the ranks, 30 doubles all-reduced per
    iteration; 10 fixed iterations are timed.
Unit: source-point evaluations per second (Mpts/s) = source points x linearisations / device time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C4_PARAMS = dict(voxel_size=0.05, max_corresp_dist=1.0, epsilon=0.005, rel_fitness=1e-7, rel_rmse=1e-7, max_iterations=100)
C5_TRUE = dict(t=(0.3, -0.2, 0.1), rpy_deg=(0.5, 0.3, -0.8))


# ------------------------------------------------------------------------------------------------ inputs
def rig_pose(i):
    rig = [(0.0, 0.0, 0.0, 0.0, 0.0, 1.8), (0.0, 0.02, 1.5, 1.0, 0.0, 1.8), (0.01, 0.0, -1.5, 1.0, -0.6, 1.8),
           (0.0, -0.02, 3.0, -0.8, 0.3, 1.9), (-0.01, 0.01, 0.6, 0.4, 0.5, 1.7)]
    return np.array(rig[i])


def pose_matrix(p):
    from multi_sensor_slam_tookit_b200 import synth
    T = np.eye(4)
    T[:3, :3] = synth.rot_zyx(p[0], p[1], p[2]); T[:3, 3] = p[3:6]
    return T


def c4_clouds(n_rings=64, n_cols=1024):
    from multi_sensor_slam_tookit_b200 import synth
    scene = synth.CityBlock()
    out = []
    for i in range(5):
        raw = synth.ring_scan(scene, rig_pose(i), n_rings=n_rings, n_cols=n_cols, elev_deg=(-22.5, 22.5),
                              seed=synth.MASTER_SEED + 40 + i, noise=0.01, dropout=0.02)
        out.append(np.stack([raw["x"], raw["y"], raw["z"]], 1).astype(np.float64))
    return out


def c4_pairs(mode="fitness"):
    """fitness mode: the 20 ordered pairs of multi_lidar_calibrator.py:302-321; standard mode: every other lidar onto the
    target lidar 0, 4 pairs (multi_lidar_calibrator.py:202-219)."""
    if mode == "standard":
        return [(s, 0) for s in range(1, 5)]
    return [(s, t) for t in range(5) for s in range(5) if s != t]        # 20 ordered pairs


def c4_init(s, t):
    truth = np.linalg.inv(pose_matrix(rig_pose(t))) @ pose_matrix(rig_pose(s))
    from multi_sensor_slam_tookit_b200 import synth
    D = np.eye(4)
    r = np.deg2rad((0.8, -0.5, 1.2))
    D[:3, :3] = synth.rot_zyx(r[0], r[1], r[2]); D[:3, 3] = (0.10, -0.06, 0.04)
    return D @ truth, truth


def c5_clouds_torch(n, device, seed=20261018 + 5):
    """n noise-free surface samples of the 4x4-tiled city block (ground, walls, roofs), then target = S + N(0, 0.02),
    source = T^-1 (S + N(0, 0.02)). Generated on the device from the seed; returned as host float64 arrays."""
    import torch
    from multi_sensor_slam_tookit_b200 import synth
    scene = synth.CityBlock(tiles=4)
    g = torch.Generator(device=device); g.manual_seed(seed)
    ext = torch.tensor(scene.box_max - scene.box_min, device=device)
    bmin = torch.tensor(scene.box_min, device=device)
    bmax = torch.tensor(scene.box_max, device=device)
    wall_area = 2 * (ext[:, 0] + ext[:, 1]) * ext[:, 2]
    roof_area = ext[:, 0] * ext[:, 1]
    ground_area = (2 * scene.half) ** 2
    areas = torch.tensor([ground_area, float(wall_area.sum()), float(roof_area.sum())], dtype=torch.float64)
    counts = (areas / areas.sum() * n).long()
    counts[0] = n - counts[1] - counts[2]
    ng, nw, nr = [int(c) for c in counts]
    U = lambda m: torch.rand(m, generator=g, device=device, dtype=torch.float64)       # noqa: E731
    pts = torch.empty((n, 3), dtype=torch.float64, device=device)
    pts[:ng, 0] = (U(ng) * 2 - 1) * scene.half; pts[:ng, 1] = (U(ng) * 2 - 1) * scene.half; pts[:ng, 2] = 0.0
    b = torch.multinomial((wall_area / wall_area.sum()).float(), nw, replacement=True, generator=g)
    ex, ey = ext[b, 0], ext[b, 1]
    per = U(nw) * 2 * (ex + ey)
    x = torch.where(per < ex, per, torch.where(per < ex + ey, ex, torch.where(per < 2 * ex + ey, 2 * ex + ey - per, torch.zeros_like(per))))
    y = torch.where(per < ex, torch.zeros_like(per), torch.where(per < ex + ey, per - ex, torch.where(per < 2 * ex + ey, ey, 2 * (ex + ey) - per)))
    pts[ng:ng + nw, 0] = bmin[b, 0] + x; pts[ng:ng + nw, 1] = bmin[b, 1] + y; pts[ng:ng + nw, 2] = U(nw) * ext[b, 2]
    b = torch.multinomial((roof_area / roof_area.sum()).float(), nr, replacement=True, generator=g)
    pts[ng + nw:, 0] = bmin[b, 0] + U(nr) * ext[b, 0]; pts[ng + nw:, 1] = bmin[b, 1] + U(nr) * ext[b, 1]; pts[ng + nw:, 2] = bmax[b, 2]
    perm = torch.randperm(n, generator=g, device=device)
    pts = pts[perm]
    tgt = pts + 0.02 * torch.randn((n, 3), generator=g, device=device, dtype=torch.float64)
    srcw = pts + 0.02 * torch.randn((n, 3), generator=g, device=device, dtype=torch.float64)
    r = np.deg2rad(C5_TRUE["rpy_deg"])
    T = np.eye(4); T[:3, :3] = synth.rot_zyx(r[0], r[1], r[2]); T[:3, 3] = C5_TRUE["t"]
    Rt = torch.tensor(T[:3, :3], device=device); tt = torch.tensor(T[:3, 3], device=device)
    src = (srcw - tt) @ Rt                                # R^T (p - t), row-vector form
    del pts, srcw, perm
    return src.cpu().numpy(), tgt.cpu().numpy(), T


# ------------------------------------------------------------------------------------------------ C4
def run_c4(rank=0, world=1, n_cols=1024, repeats=1, mode="fitness"):
    """Each rank calibrates its round-robin share of the 20 ordered pairs, host buffers in, result out (what
    Calibration.compute_gicp_transformation does per pair). Returns per-rank totals."""
    from multi_sensor_slam_tookit_b200 import gicp
    clouds = c4_clouds(n_cols=n_cols)
    P = C4_PARAMS
    pairs = c4_pairs(mode)
    mine = [i for i in range(len(pairs)) if gicp.pair_owner(i, world) == rank]
    stats = dict(pairs=len(mine), src_evals=0.0, gpu_ms=0.0, wall_s=0.0, iters=0, max_t_err=0.0, max_r_err=0.0, launches=0,
                 pts_in=0, pts_ds=0, prep_ms=0.0)
    for rep in range(repeats + 1):                       # first pass is the warm-up
        acc = dict(stats)
        for i in mine:
            s, t = pairs[i]
            init, truth = c4_init(s, t)
            t0 = time.perf_counter()
            sp = gicp.PointCloud(clouds[s]).voxel_down_sample(P["voxel_size"])
            tp = gicp.PointCloud(clouds[t]).voxel_down_sample(P["voxel_size"])
            sp.estimate_normals(); tp.estimate_normals()
            prep = sp.lastGpuMs() + tp.lastGpuMs()
            res = gicp.registration_generalized_icp(sp, tp, P["max_corresp_dist"], init,
                                                    gicp.TransformationEstimationForGeneralizedICP(P["epsilon"]),
                                                    gicp.ICPConvergenceCriteria(P["rel_fitness"], P["rel_rmse"], P["max_iterations"]))
            acc["wall_s"] += time.perf_counter() - t0
            acc["gpu_ms"] += res.gpu_ms
            acc["prep_ms"] += prep
            acc["iters"] += res.iterations
            acc["launches"] += res.gpu_launches
            acc["src_evals"] += float(len(sp)) * (res.iterations + 1)
            acc["pts_in"] += len(clouds[s]); acc["pts_ds"] += len(sp)
            dT = np.linalg.inv(truth) @ res.transformation
            acc["max_t_err"] = max(acc["max_t_err"], float(np.linalg.norm(dT[:3, 3])))
            acc["max_r_err"] = max(acc["max_r_err"], float(np.arccos(np.clip((np.trace(dT[:3, :3]) - 1) / 2, -1, 1))))
        out = acc
    return out


def cpu_c4_pair(threads, s=1, t=0, n_cols=1024):
    """The same pair on the CPU oracle (bounded sample for cpu_baseline): returns (src_evals, seconds_total, seconds_register)."""
    from oracle import pyoracle as O
    clouds = c4_clouds(n_cols=n_cols)
    P = C4_PARAMS
    init, _ = c4_init(s, t)
    t0 = time.perf_counter()
    sp, _ = O.o3d_voxel_down_sample(clouds[s], P["voxel_size"])
    tp, _ = O.o3d_voxel_down_sample(clouds[t], P["voxel_size"])
    _, sc = O.gicp_normals_covs(sp, 30, P["epsilon"], threads)
    _, tc = O.gicp_normals_covs(tp, 30, P["epsilon"], threads)
    t1 = time.perf_counter()
    r = O.GicpOracle(sp, sc, tp, tc, threads).register(init, P["max_corresp_dist"], P["rel_fitness"], P["rel_rmse"], P["max_iterations"])
    t2 = time.perf_counter()
    return float(len(sp)) * (r["iterations"] + 1), t2 - t0, t2 - t1, r["iterations"]


# ------------------------------------------------------------------------------------------------ C5
_KEEP = []


def run_c5(n_points, rank=0, world=1, comm=None, iterations=10, repeats=2, device="cuda"):
    from multi_sensor_slam_tookit_b200 import gicp
    t0 = time.perf_counter()
    src, tgt, T_true = c5_clouds_torch(n_points, device)
    t_gen = time.perf_counter() - t0
    if comm is not None:
        # NCCL sets its channels up at the first collective of each kind on a communicator (hundreds of ms): not part of the step
        comm.allreduce(np.zeros(4))
        warm = gicp.PointCloud.from_host_sharded(np.zeros((4096, 3)) + np.arange(4096)[:, None], comm)
        del warm
        if not os.environ.get("B2_GICP_NO_FUSED_EXCHANGE"):
            # likewise the first mapping of a peer GPU's memory in this process (cudaIpcOpenMemHandle enables peer access: ~13 ms per peer)
            wg = gicp.GeneralizedICP(1.0, 0.005); wg.setShard(comm); wg.setupExchange()
            _KEEP.append(wg)      # kept for the life of the process: its peers have the area mapped, it is not freed under them
    t0 = t_e2e = time.perf_counter()
    if comm is not None:
        # sharded set-up: 1/world of each cloud over this rank's PCIe link + all-gather, 1/world of the 30-NN normals + all-gather,
        # the source grid over this rank's rows only (the target and its two grids are replicated)
        sp, tp = gicp.PointCloud.from_host_sharded(src, comm), gicp.PointCloud.from_host_sharded(tgt, comm)
    else:
        sp, tp = gicp.PointCloud(src), gicp.PointCloud(tgt)
    n_src = len(src)
    del src, tgt
    t_up = time.perf_counter() - t0
    if comm is not None:
        sp.estimate_normals_sharded(comm); n_ms = sp.lastGpuMs()
        tp.estimate_normals_sharded(comm); n_ms += tp.lastGpuMs()
    else:
        sp.estimate_normals(); n_ms = sp.lastGpuMs()
        tp.estimate_normals(); n_ms += tp.lastGpuMs()
    g = gicp.GeneralizedICP(1.0, 0.005, -1.0, -1.0, iterations)         # negative thresholds: never "converged", fixed iterations
    t0 = time.perf_counter()
    g.setInputTarget(tp)
    if comm is not None:
        g.setInputSourceBlocks(sp, rank, world)
    else:
        g.setInputSource(sp)
    t_index = time.perf_counter() - t0
    fused = False
    t0 = time.perf_counter()
    if comm is not None:
        g.setShard(comm)
        if not os.environ.get("B2_GICP_NO_FUSED_EXCHANGE"):
            # fused linearise + exchange: the CUDA IPC handles of the ranks' exchange areas travel over the communicator
            fused = g.setupExchange()
    t_peers = time.perf_counter() - t0
    info = g.indexInfo()
    runs = []
    e2e_s = None
    for _ in range(repeats + 1):
        res = g.align(np.eye(4), want_correspondences=False)
        if e2e_s is None:
            e2e_s = time.perf_counter() - t_e2e            # upload + normals + index builds + the first align
        runs.append(res)
    runs.sort(key=lambda r: r.gpu_ms)
    res = runs[len(runs) // 2]                            # median of the aligns (3 by default)
    dT = np.linalg.inv(T_true) @ res.transformation
    return dict(points=int(n_points), iterations=int(res.iterations), evaluations=int(res.iterations + 1), gpu_ms=float(res.gpu_ms),
                launches=int(res.gpu_launches), fitness=float(res.fitness), inlier_rmse=float(res.inlier_rmse),
                t_err=float(np.linalg.norm(dT[:3, 3])), r_err=float(np.arccos(np.clip((np.trace(dT[:3, :3]) - 1) / 2, -1, 1))),
                evaluation_ms=[round(float(v), 3) for v in res.evaluation_ms], shard_points=int(info["shard_points"]), cell_edge=float(info["target_cell_edge"]), ppc=float(info["target_points_per_cell"]),
                e2e_s=e2e_s, setup_s=dict(generate=t_gen, upload=t_up, normals_ms=n_ms, index=t_index, exchange_setup=t_peers), fused_exchange=bool(fused))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c5-points", type=int, default=50_000_000)
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="also time one C4 pair on the CPU oracle")
    ap.add_argument("--c4-cols", type=int, default=1024)
    args = ap.parse_args()
    import torch
    from multi_sensor_slam_tookit_b200 import capi, gicp
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    capi.check(capi.lib().b2_set_device(local))
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [gicp.Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = gicp.Communicator(ids[0], rank, world)
    out = {"world": world}
    if not args.skip_c4:
        r = run_c4(rank, world, n_cols=args.c4_cols)
        r["mpts_per_s"] = r["src_evals"] / max(r["gpu_ms"], 1e-9) / 1e3
        out["c4"] = r
    if not args.skip_c5:
        r = run_c5(args.c5_points, rank, world, comm)
        r["mpts_per_s"] = r["points"] * r["evaluations"] / r["gpu_ms"] / 1e3
        out["c5"] = r
    if args.cpu and rank == 0:
        ev, tt, tr, it = cpu_c4_pair(os.cpu_count() or 1, n_cols=args.c4_cols)
        out["cpu_c4_pair"] = dict(src_evals=ev, total_s=tt, register_s=tr, iterations=it, mpts_per_s=ev / tr / 1e6, cores=os.cpu_count())
    print(f"rank {rank}: " + json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
