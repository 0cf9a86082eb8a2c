#!/usr/bin/env python
"""bench.py — scan-to-map LM throughput on the C1 workload of BASELINE.json (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

One JSON line on stdout (rank 0). A "step" is what the reference does per scan in scan2MapOptimization
(liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:1282-1310): index build of the two map clouds (kdtree->setInputCloud, :1289-1290)
+ the LM loop, one VLP-16 scan against the 100 000-point map. BOTH arms time this step (config.step is identical).
  value      LM iterations/s of that step with map points and scan features already resident in HBM (device time, CUDA
             events on the library's streams from the start of the index build to the end of the solve, L2 flushed between steps)
  solve_only the LM loop alone on a resident index (pairs with cpu_baseline.solve_only)
  e2e        the same step through the C ABI with HOST buffers: set_map + set_scan + solve per step, wall clock
  roofline   k_s2m_iteration: algorithmic bytes (72 B + 16 B per candidate in the 27-cell block, per feature) over
             the measured launch time, against MEASURED_PEAKS.json's HBM copy bandwidth
  batched    B scans (pose hypotheses) against the one map in a single launch per iteration
  cpu_baseline  the CPU oracle (restatement of the reference path) on the box's host cores
--impl reference times that CPU path alone (PCL/OpenCV are not installable here: the oracle is the reference arm).
  registration  front_end_c2: the per-scan front end on the 128 x 1024 scan of config C2; then the NDT / GICP half of the metric (Mpts/s): C3 NDT pair, C4 batched GICP pairs (round-robin over ranks),
             C5 50 M-point map-to-map GICP (source sharded over ranks, NCCL all-reduce of 30 doubles per iteration)
Multi-GPU: C1 does not shard — N GPUs run N independent replicas, no collective on the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scan_to_map_lm_iters_per_sec"
UNIT = "iters/s"
# the step both arms time (the driver compares the two lines; they must describe the same work)
STEP = "index build of both map clouds (kdtree setInputCloud x2, mapOptmization.cpp:1289-1290) + scan2MapOptimization loop (<=30 LM iterations, stops on convergence)"


def load_c1():
    d = np.load(os.path.join(ROOT, "tests", "golden", "c1_input.npz"))
    return {k: d[k] for k in d.files}


def workload_name(c1):
    return (f"C1 LIO-SAM scan-to-map LM: VLP-16 16x1800 scan ({len(c1['scan_corner'])} corner + {len(c1['scan_surf'])} surf "
            f"features) vs {len(c1['map_corner']) + len(c1['map_surf'])}-pt map ({len(c1['map_corner'])} corner)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return int(t[kernel]["dram_bytes_per_launch_last"])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ candidates per feature
def pose_matrix(pose):
    r, p, y = [float(v) for v in pose[:3]]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, sy * sr + cy * sp * cr],
                  [sy * cp, cy * cr + sy * sp * sr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    return R, np.asarray(pose[3:6], np.float64)


def candidates_per_feature(map_pts, scan_pts, pose, h=1.0078125):
    """Points inside the 3x3x3 cell block the kernel scans for each feature (pure numpy, same grid geometry)."""
    if len(map_pts) == 0 or len(scan_pts) == 0:
        return np.zeros(len(scan_pts))
    R, t = pose_matrix(pose)
    q = scan_pts[:, :3].astype(np.float64) @ R.T + t
    o = map_pts[:, :3].min(0).astype(np.float64)
    mc = np.floor((map_pts[:, :3] - o) / h).astype(np.int64)
    dims = mc.max(0) + 2
    cnt = np.zeros(tuple(dims + 2), np.int64)          # +1 halo on each side
    np.add.at(cnt, (mc[:, 0] + 1, mc[:, 1] + 1, mc[:, 2] + 1), 1)
    qc = np.floor((q - o) / h).astype(np.int64)
    ok = np.all((qc >= -1) & (qc <= dims - 1 + 1), axis=1)
    qc = np.clip(qc, -1, dims) + 1
    tot = np.zeros(len(q), np.int64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                ix = np.clip(qc[:, 0] + dx, 0, cnt.shape[0] - 1)
                iy = np.clip(qc[:, 1] + dy, 0, cnt.shape[1] - 1)
                iz = np.clip(qc[:, 2] + dz, 0, cnt.shape[2] - 1)
                tot += cnt[ix, iy, iz]
    return np.where(ok, tot, 0)


def algorithmic_bytes_per_iteration(c1, pose):
    cc = candidates_per_feature(c1["map_corner"], c1["scan_corner"], pose)
    cs = candidates_per_feature(c1["map_surf"], c1["scan_surf"], pose)
    nfeat = len(cc) + len(cs)
    cand = float(cc.sum() + cs.sum())
    return 72.0 * nfeat + 16.0 * cand + 28 * 8, cand / max(nfeat, 1), 72.0 * nfeat + 28 * 8


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        try:
            self.proc.terminate()
            self.t.join(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def host_cores():
    """Threads for the CPU arm: the cores this process may run on, minus one per other rank of the job (while rank 0 times the
    CPU path the other ranks sit in a collective, each spinning on one core; an OpenMP team that is one thread wider than the
    free cores collapses — SCALE_r01 showed 67-89 iters/s instead of 2 300). torchrun exports OMP_NUM_THREADS=1, but the
    oracle's parallel regions carry an explicit num_threads clause; the team size really obtained is what gets reported."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, n - (int(os.environ.get("WORLD_SIZE", "1")) - 1))


def cpu_solve_rate(c1, threads, budget_s, include_build):
    """iters/s of the CPU restatement: solve only (trees built) or set_map + set_scan + solve per step."""
    from oracle import pyoracle as O
    s = O.Scan2Map(threads)
    s.set_map(c1["map_corner"], c1["map_surf"])
    s.set_scan(c1["scan_corner"], c1["scan_surf"])
    for _ in range(2):
        s.solve(c1["pose_guess"])
    iters, steps, t0 = 0, 0, time.perf_counter()
    while True:
        if include_build:
            s.set_map(c1["map_corner"], c1["map_surf"])
            s.set_scan(c1["scan_corner"], c1["scan_surf"])
        r = s.solve(c1["pose_guess"])
        iters += r["iters"]; steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or steps >= 100000:
            break
    return iters / el, steps, el, r["iters"]


def run_reference(args, rank, world):
    if rank != 0:
        return
    c1 = load_c1()
    cores = host_cores()
    t_all = time.perf_counter()
    # warm-up and timed steps: one step = kd-tree build of both maps + the LM loop (what the reference does per scan)
    from oracle import pyoracle as O
    s = O.Scan2Map(cores)
    def step():
        s.set_map(c1["map_corner"], c1["map_surf"])
        s.set_scan(c1["scan_corner"], c1["scan_surf"])
        return s.solve(c1["pose_guess"])["iters"]
    for _ in range(max(args.warmup, 3)):
        step()
    iters, per_step, t0 = 0, [], time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        iters += step()
        per_step.append(time.perf_counter() - ts)
    el = time.perf_counter() - t0
    val = iters / el
    used = O.omp_threads(cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(c1), "step": STEP, "iters_per_step": iters / args.steps,
                       "median_ms_per_step": 1e3 * float(np.median(per_step))},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{args.steps} full steps (index build + LM loop) of the same workload, OpenMP over features "
                                       f"as mapOptmization.cpp:978,1070, {used} threads (the reference ships numberOfCores: 4)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from multi_sensor_slam_tookit_b200 import capi
    from multi_sensor_slam_tookit_b200.registration import ScanToMapOptimizer
    L = capi.lib()
    if L.b2_device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    capi.check(L.b2_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    c1 = load_c1()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()      # noqa: E731
    mc, ms, sc, ss = pin(c1["map_corner"]), pin(c1["map_surf"]), pin(c1["scan_corner"]), pin(c1["scan_surf"])
    guess = c1["pose_guess"].copy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")                     # > 126 MB L2

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    g = ScanToMapOptimizer()
    g.setInputMap(mc, ms)
    g.setInputScan(sc, ss)

    def solve(max_it=30):
        g.transformTobeMapped = guess.copy()
        return g.scan2MapOptimization(max_it, want_matP=False)

    r0 = solve()
    iters_needed = r0["iters"]
    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        flush_l2(); solve()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- value: map points and scan features resident in HBM; one step = index build of both maps + LM loop (the reference's
    # per-scan step); device time from the start of the build to the end of the solve (events on the library's streams)
    for _ in range(W):
        flush_l2(); g.rebuildMapIndex(); solve()
    barrier()
    l0 = L.b2_kernel_launch_count()
    dev_ms, wall_ms, iters, solve_ms = 0.0, 0.0, 0, 0.0
    for _ in range(K):
        flush_l2()
        t0 = time.perf_counter()
        g.rebuildMapIndex()
        r = solve()
        wall_ms += (time.perf_counter() - t0) * 1e3
        dev_ms += g.lastStepGpuMs()
        solve_ms += g.lastGpuMs()[0]
        iters += r["iters"]
    launches = L.b2_kernel_launch_count() - l0
    barrier()
    # ---- roofline: the LM loop cut at the iterations it needs, so every launch in the timed span is an active one
    act_ms = 0.0
    for _ in range(K):
        flush_l2(); solve(iters_needed)
        act_ms += g.lastGpuMs()[0]
    barrier()
    # ---- e2e: host buffers through the C ABI every step (set_map + set_scan + solve), wall clock
    for _ in range(W):
        g.setInputMap(mc, ms); g.setInputScan(sc, ss); solve()
    barrier()
    e_iters, t0 = 0, time.perf_counter()
    for _ in range(K):
        g.setInputMap(mc, ms); g.setInputScan(sc, ss)
        e_iters += solve()["iters"]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # ---- local map on the device: key frames resident in HBM, extractCloud + index + solve per step
    from multi_sensor_slam_tookit_b200 import synth
    from multi_sensor_slam_tookit_b200.registration import LocalMap
    # 50 overlapping key frames (consecutive scans see the same surfaces): ~1 M points into the two map VoxelGrids, ~100 k out
    kfs = synth.keyframes_overlapping(c1["map_corner"], c1["map_surf"], 50, 0.2)
    lm = LocalMap(0.2, 0.4, surroundingKeyframeSearchRadius=1e9)
    for kc, ks_, kp in kfs:
        lm.saveKeyFrame(kc, ks_, kp)
    order = list(range(len(kfs)))

    def lm_step():
        lm.extractCloud(order); g.setInputMapFromLocalMap(lm); g.setInputScan(sc, ss)
        return solve()["iters"]
    for _ in range(W):
        lm_step()
    barrier()
    lm_iters, lm_ms, t0 = 0, 0.0, time.perf_counter()
    for _ in range(K):
        lm_iters += lm_step(); lm_ms += lm.lastGpuMs()[0]
    torch.cuda.synchronize()
    lm_s = time.perf_counter() - t0
    lm_pts = (int(sum(len(k[0]) for k in kfs)), int(sum(len(k[1]) for k in kfs)), int(len(lm.get("cornerDS"))), int(len(lm.get("surfDS"))))
    g.setInputMap(mc, ms)
    barrier()
    # ---- batched: B (scan, initial guess) problems against the resident map, two launches per LM iteration (search, fits).
    # The scans are DISTINCT: 16 VLP-16 sweeps ray-cast at 16 poses 0.6 m apart along the street, features from the library's
    # own front end (b2_scan_*) and VoxelGrids, each registered from 16 different initial guesses.
    batched = None
    B = args.batch
    if B > 1:
        from multi_sensor_slam_tookit_b200.frontend import ScanFrontEnd
        from multi_sensor_slam_tookit_b200.registration import VoxelGrid
        n_distinct = min(16, B)
        scene = synth.CityBlock(synth.MASTER_SEED)
        fe = ScanFrontEnd()
        vgc, vgs = VoxelGrid(), VoxelGrid()
        vgc.setLeafSize(0.2, 0.2, 0.2); vgs.setLeafSize(0.4, 0.4, 0.4)
        scans, truths = [], []
        for k in range(n_distinct):
            pk = c1["pose_truth"].astype(np.float64).copy()
            pk[3] += (k - n_distinct // 2) * 0.6 + 0.3
            raw = synth.ring_scan(scene, pk, seed=synth.MASTER_SEED + 7000 + k)
            fe.projectPointCloud(raw, imu=None, deskew=False)
            f = fe.extractFeatures()
            vgc.setInputCloud(f["corner"]); vgs.setInputCloud(f["surf"])
            scans.append((np.ascontiguousarray(vgc.filter()), np.ascontiguousarray(vgs.filter())))
            truths.append(pk.astype(np.float32))
        rng = np.random.default_rng(7)
        idx = np.arange(B) % n_distinct
        poses = np.stack([truths[i] for i in idx]).astype(np.float32)
        poses[:, 3:] += rng.uniform(-0.15, 0.15, (B, 3)).astype(np.float32)
        poses[:, :3] += np.deg2rad(rng.uniform(-1.0, 1.0, (B, 3))).astype(np.float32)
        gb = ScanToMapOptimizer(max_batch=B)
        gb.setInputMap(mc, ms)
        gb.setInputScanBatch([scans[i][0] for i in idx], [scans[i][1] for i in idx])
        rb = gb.scan2MapOptimizationBatch(poses)
        b_max = int(rb["iters"].max())
        # one counted solve: candidates the search really loads, and the features of every active (scan, iteration) pair
        gb.countCandidates(True)
        rc = gb.scan2MapOptimizationBatch(poses, b_max)
        cand_total = gb.countCandidates(False)
        feat_per_scan = np.array([len(scans[i][0]) + len(scans[i][1]) for i in idx], np.float64)
        feat_iters = float((feat_per_scan * rc["iters"]).sum())
        for _ in range(3):
            flush_l2(); gb.scan2MapOptimizationBatch(poses, b_max)
        Kb = max(3, min(K, 20))
        b_ms, b_iters = 0.0, 0
        for _ in range(Kb):
            flush_l2()
            rb = gb.scan2MapOptimizationBatch(poses, b_max)
            b_ms += gb.lastGpuMs()[0]
            b_iters += int(rb["iters"].sum())
        b_bytes = 72.0 * feat_iters + 16.0 * float(cand_total) + 224.0 * float(rc["iters"].sum())
        batched = {"scans": B, "distinct_scans": n_distinct, "steps": Kb, "value": b_iters / (b_ms * 1e-3), "unit": UNIT, "ms_per_step": b_ms / Kb,
                   "iters_per_step": b_iters / Kb, "converged": int(rb["converged"].sum()), "max_iters_in_batch": b_max,
                   "features_per_scan": float(feat_per_scan.mean()), "candidates_read_per_feature": float(cand_total) / max(feat_iters, 1.0),
                   "bytes_per_step": b_bytes, "gpu_launches_per_step": int(gb.lastGpuMs()[1])}
        del gb, fe, vgc, vgs
    barrier()
    clocks = sampler.stop() if sampler else None

    # ---- reductions over ranks: units summed, time = max
    t = torch.tensor([dev_ms, e2e_s, act_ms, wall_ms, solve_ms], dtype=torch.float64, device="cuda")
    u = torch.tensor([iters, e_iters], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dev_ms_m, e2e_s_m, act_ms_m, wall_ms_m, solve_ms_m = [float(v) for v in t.tolist()]
    iters_all, e_iters_all = [float(v) for v in u.tolist()]
    del g, flush
    L.b2_trim_memory()
    registration = None if args.skip_registration else run_registration(args, rank, local_rank, world, dist, torch)
    if rank == 0:
        peak, peak_src = peaks()
        abytes, cand, compulsory = algorithmic_bytes_per_iteration(c1, guess)
        launch_ms = act_ms_m / (K * iters_needed)                 # span / active launches (includes launch gaps + prepare)
        achieved = abytes / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": iters_all / (dev_ms_m * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_m / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(c1), "step": STEP,
                       "iters_per_step": iters / K, "l2": "flushed between steps (256 MiB device write)",
                       "parallelism": f"replicas x{world} (C1 does not shard; no data-path collective)",
                       "timing": "device ms per step = CUDA events on the library's streams, start of the index build -> end of the solve",
                       "wall_ms_per_step": wall_ms_m / K},
            "solve_only": {"value": iters_all / (solve_ms_m * 1e-3), "unit": UNIT, "ms_per_step": solve_ms_m / K,
                           "step": "scan2MapOptimization loop on a resident index (pairs with cpu_baseline.solve_only)"},
            "e2e": {"value": e_iters_all / e2e_s_m, "unit": UNIT,
                    "h2d_bytes_per_step": int(mc.nbytes + ms.nbytes + sc.nbytes + ss.nbytes + 592 + 16),
                    "d2h_bytes_per_step": int(592 + 4 + 2 * 24), "ms_per_step": 1e3 * e2e_s_m / K,
                    "step": "the same step from HOST buffers: b2_s2m_set_map (upload + index build) + b2_s2m_set_scan + b2_s2m_solve, pinned memory, wall clock"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_s2m_iteration", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic("k_s2m_iteration"), "peak_source": peak_src,
                         "bytes_per_launch": abytes, "compulsory_bytes_per_launch": compulsory,
                         "candidates_per_feature": cand, "launch_ms": launch_ms,
                         "note": "span of a solve capped at the iterations it needs / launches; working set (1.7 MB) is L2-resident "
                                 "after first touch, so this is L2+DRAM bytes over time, see DESIGN.md"},
            "clocks": clocks,
        }
        from oracle import pyoracle as O

        def cpu_lm_step():
            cat_c = np.concatenate([O.transform_cloud(k[0], k[2], cores) for k in kfs]); cat_s = np.concatenate([O.transform_cloud(k[1], k[2], cores) for k in kfs])
            dc, ds_ = O.voxel_grid(cat_c, 0.2)["out"], O.voxel_grid(cat_s, 0.4)["out"]
            so.set_map(dc, ds_); so.set_scan(c1["scan_corner"], c1["scan_surf"])
            return so.solve(c1["pose_guess"])["iters"]
        cores = host_cores()
        so = O.Scan2Map(cores)
        cpu_lm_step()
        ci, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 4.0:
            ci += cpu_lm_step()
        cpu_lm = ci / (time.perf_counter() - t0)
        # the same step with the assembled map crossing PCIe instead (what a caller without the device-resident key frames pays)
        g_up = ScanToMapOptimizer()
        mcd, msd = pin(lm.get("cornerDS")), pin(lm.get("surfDS"))
        for _ in range(W):
            g_up.setInputMap(mcd, msd); g_up.setInputScan(sc, ss); g_up.transformTobeMapped = guess.copy(); g_up.scan2MapOptimization(30, want_matP=False)
        ui, t0 = 0, time.perf_counter()
        for _ in range(K):
            g_up.setInputMap(mcd, msd); g_up.setInputScan(sc, ss); g_up.transformTobeMapped = guess.copy()
            ui += g_up.scan2MapOptimization(30, want_matP=False)["iters"]
        up_s = time.perf_counter() - t0
        del g_up
        line["local_map"] = {
            "workload": f"extractCloud over {len(kfs)} overlapping device-resident key frames ({lm_pts[0]} corner + {lm_pts[1]} surf points -> {lm_pts[2]} + {lm_pts[3]} "
                        "after VoxelGrid 0.2 / 0.4), index build, set_scan, LM loop; the map never crosses PCIe (mapOptmization.cpp:899-938, 1282-1310)",
            "e2e": {"value": lm_iters / lm_s, "unit": UNIT, "ms_per_step": 1e3 * lm_s / K, "h2d_bytes_per_step": int(sc.nbytes + ss.nbytes + 592 + 16 + 96 * len(kfs)),
                    "d2h_bytes_per_step": 644 + 2 * 28}, "extract_gpu_ms": lm_ms / K,
            "upload_path": {"value": ui / up_s, "unit": UNIT, "ms_per_step": 1e3 * up_s / K,
                            "step": "the already assembled + downsampled map uploaded from the host (set_map) + set_scan + LM loop: what the "
                                    "device-resident path has to beat although it also does the assembly and both VoxelGrids"},
            "cpu_baseline": {"value": cpu_lm, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "the same step on the oracle (transformPointCloud x 24, two VoxelGrids, two kd-trees, LM loop) for 4 s"}}
        if batched:
            bb = batched["bytes_per_step"] / (batched["ms_per_step"] * 1e-3) / 1e9
            batched["roofline"] = {"kernel": "k_s2m_iteration<1,1,1,1> (search) + <1,1,1,2> (fits)", "bound": "hbm", "achieved": bb, "peak": peak,
                                   "unit": "GB/s", "frac": bb / peak, "traffic": measured_traffic("k_s2m_batched"),
                                   "note": "algorithmic bytes = 72 B per feature + 16 B per candidate the search LOADED (counted on the device in a "
                                           "separate solve) + 224 B per (scan, iteration), over the device span of the step. The map (1.6 MB) is "
                                           "L2-resident: this is L2-to-SM traffic against the HBM peak, DRAM itself stays nearly idle (see profiles/)"}
            line["batched"] = batched
        # CPU baseline on this box's host cores, bounded samples of the same step
        cores = host_cores()
        used = O.omp_threads(cores)
        v_full, st2, el2, it1 = cpu_solve_rate(c1, cores, 6.0, True)
        v_solve, st1, el1, _ = cpu_solve_rate(c1, cores, 6.0, False)
        v_4, _, _, _ = cpu_solve_rate(c1, min(4, cores), 3.0, True)
        line["cpu_baseline"] = {"value": v_full, "unit": UNIT, "cores": used, "kind": "port",
                                "sample": f"{st2} steps (kd-tree build of both maps + LM loop, {it1} iterations each) of the same scan/map in {el2:.1f} s",
                                "solve_only": v_solve, "solve_only_sample": f"{st1} solves in {el1:.1f} s, kd-trees prebuilt",
                                "threads4": v_4, "threads4_note": "the same step at the reference's shipped numberOfCores: 4 (config/params.yaml:72)",
                                "note": "CPU restatement of the reference path (oracle/), OpenMP over features as the reference"}
        if registration:
            line["registration"] = registration
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ NDT / GICP workloads
def run_registration(args, rank, local_rank, world, dist, torch):
    """Configs C3 (NDT), C4 (batched GICP pairs) and C5 (sharded map-to-map GICP) of BASELINE.json — the "NDT/GICP Mpts/s" half
    of the metric. Unit everywhere: source-point evaluations per second. Returns the dict rank 0 prints (None elsewhere)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gicp_bench as GB
    import ndt_bench as NB
    from multi_sensor_slam_tookit_b200 import gicp
    peak, peak_src = peaks()
    out = {}

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- C2: the per-scan front end does not shard either -> rank 0 only
    if rank == 0:
        import frontend_bench as FB
        inp = FB.c2_inputs()
        r = FB.run_c2(inputs=inp)
        cpu = FB.cpu_c2(inp, 5.0)
        ach = r["algorithmic_bytes"] / ((r["project_gpu_ms"] + r["features_gpu_ms"]) * 1e-3) / 1e9
        out["front_end_c2"] = {
            "workload": f"C2 LIO-SAM imageProjection + deskew + featureExtraction: 128 x 1024 scan, {r['points_in']} returns -> {r['extracted']} extracted, "
                        f"{r['corner']} corner + {r['surf']} surf features", "value": r["mpts_per_s"], "unit": "Mpts/s",
            "project_gpu_ms": r["project_gpu_ms"], "features_gpu_ms": r["features_gpu_ms"],
            "e2e": {"value": r["e2e_mpts_per_s"], "unit": "Mpts/s", "ms": r["e2e_ms"], "h2d_bytes": int(32 * r["points_in"] + 4 * 8 * 62),
                    "d2h_bytes": int(24 * r["extracted"] + 16 * (r["corner"] + r["surf"]) + 1024),
                    "step": "b2_scan_project + b2_scan_extract_features, PointXYZIRT records in, cloud_info arrays + corner / surf clouds out"},
            "roofline": {"kernel": "k_scan_* (10 launches)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": measured_traffic("k_scan_all"), "peak_source": peak_src, "bytes_per_launch": r["algorithmic_bytes"],
                         "note": "52 B per input point + 20 B per extracted point (SURVEY.md 8d) over the device span of both calls; 4.2 MB in, "
                                 "launch-latency bound at this size (one scan)"},
            "cpu_baseline": {"value": cpu["mpts_per_s"], "unit": "Mpts/s", "cores": 1, "kind": "port",
                             "sample": f"{cpu['reps']} scans: project {cpu['project_ms']:.2f} ms + features {cpu['features_ms']:.2f} ms (serial, as the reference's two nodes are)"}}
    # ---- pcl::VoxelGrid at the sizes the local map feeds it (0.3-1 M points in, ~100 k out; 4 M = a long session)
    if rank == 0:
        from multi_sensor_slam_tookit_b200.registration import VoxelGrid
        from oracle import pyoracle as O
        c1v = load_c1()
        rngv = np.random.default_rng(3)
        out["voxel_grid"] = {}
        for n_in in (1_000_000, 4_000_000):
            base = c1v["map_surf"][rngv.integers(0, len(c1v["map_surf"]), n_in)].copy()
            base[:, :3] += rngv.normal(0.0, 0.05, (n_in, 3)).astype(np.float32)
            hostbuf = torch.from_numpy(np.ascontiguousarray(base)).pin_memory().numpy()
            vg = VoxelGrid(); vg.setLeafSize(0.4, 0.4, 0.4); vg.setInputCloud(hostbuf)
            for _ in range(3):
                res = vg.filter()
            reps, dev, t0 = 10, 0.0, time.perf_counter()
            for _ in range(reps):
                res = vg.filter(); dev += vg.lastGpuMs()
            wall = time.perf_counter() - t0
            t0 = time.perf_counter(); ref = O.voxel_grid(base, 0.4)["out"]; cpu_s = time.perf_counter() - t0
            assert len(ref) == len(res)
            abytes = 16.0 * n_in + 16.0 * len(res)
            ach = abytes / (dev / reps * 1e-3) / 1e9
            out["voxel_grid"][f"{n_in // 1_000_000}M"] = {
                "workload": f"pcl::VoxelGrid(0.4 m) on {n_in} PointXYZI (map points resampled with 5 cm noise) -> {len(res)} voxels (mapOptmization.cpp:928,932)",
                "value": n_in / (dev / reps * 1e-3) / 1e6, "unit": "Mpts/s", "gpu_ms": dev / reps,
                "e2e": {"value": n_in * reps / wall / 1e6, "unit": "Mpts/s", "ms": 1e3 * wall / reps, "h2d_bytes": int(16 * n_in), "d2h_bytes": int(16 * len(res)),
                        "step": "b2_voxel_filter: pinned host cloud in, downsampled cloud out"},
                "roofline": {"kernel": "k_vx_* + k_rs_* (bbox, key, radix sort of (voxel, index) pairs, run heads, centroids)", "bound": "hbm", "achieved": ach,
                             "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": measured_traffic("k_vx_all") if n_in == 1_000_000 else None, "peak_source": peak_src, "bytes_per_launch": abytes,
                             "note": "16 B read + 16 B x (voxels / points) written per input point (SURVEY.md 8d) over the device span of the whole filter"},
                "cpu_baseline": {"value": n_in / cpu_s / 1e6, "unit": "Mpts/s", "cores": 1, "kind": "port", "sample": f"one filter of the same cloud in {cpu_s:.2f} s (serial, as pcl::VoxelGrid)"}}
            del vg
    # ---- C3: one NDT pair does not shard -> rank 0 only
    if rank == 0:
        inputs = NB.c3_inputs()
        r, src = NB.run_c3(inputs=inputs)
        bytes_pass = 12.0 * r["source_points"] + 96.0 * r["pairs_per_pass"]          # point + (mean, inverse covariance) per pair
        ach = bytes_pass * r["evaluations"] / (r["align_gpu_ms"] * 1e-3) / 1e9
        cpu = NB.cpu_c3(inputs, src)
        out["ndt_c3"] = {
            "workload": f"C3 multi_lidar NDT: parent {r['parent_points']} pts -> {r['voxels']} voxels of 1 m, child {r['child_points']} pts "
                        f"-> {r['source_points']} after VoxelGrid(0.1)", "value": r["mpts_per_s"], "unit": "Mpts/s",
            "e2e": {"value": r["e2e_mpts_per_s"], "unit": "Mpts/s", "ms": r["e2e_wall_ms"],
                    "step": "new handle + setInputSource + setInputTarget + align + getFitnessScore from host buffers",
                    "h2d_bytes": int(12 * (r["parent_points"] + r["source_points"])), "d2h_bytes": int(29 * 8 * r["evaluations"] + 64 + 16)},
            "iterations": r["iterations"], "evaluations": r["evaluations"], "align_gpu_ms": r["align_gpu_ms"], "gpu_launches": r["launches"],
            "t_err_m": r["t_err"], "r_err_rad": r["r_err"],
            "roofline": {"kernel": "k_ndt_derivatives", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": measured_traffic("k_ndt_derivatives"), "peak_source": peak_src, "bytes_per_launch": bytes_pass,
                         "note": "12 B point + 96 B (mean + inverse covariance) per (point, voxel) pair; 0.3 MB of voxel data, L2-resident; "
                                 "span of align / derivative passes, includes the host line search between passes"},
            "cpu_baseline": {"value": cpu["mpts_per_s"], "unit": "Mpts/s", "cores": 1, "kind": "port",
                             "sample": f"the same registration, {cpu['evaluations']} derivative passes in {cpu['align_s']:.2f} s; serial like PCL's NDT"}}
    # ---- C4: independent pairs, pair i on rank i % world, host buffers in -> transformation out
    c4 = GB.run_c4(rank, world)
    wall = allmax(c4["wall_s"]); gpu_ms = allmax(c4["gpu_ms"]); evals = allsum(c4["src_evals"]); npairs = allsum(c4["pairs"])
    iters = allsum(c4["iters"]); launches = allsum(c4["launches"]); terr = allmax(c4["max_t_err"]); rerr = allmax(c4["max_r_err"])
    if rank == 0:
        ev, tt, tr, it = GB.cpu_c4_pair(host_cores())
        ach = 144.0 * evals / (gpu_ms * 1e-3) / 1e9 / world
        out["gicp_c4"] = {
            "workload": f"C4 Multi_LiCa GICP: 5 lidars 64x1024, {int(npairs)} ordered pairs, voxel 0.05, max_corr 1.0, eps 0.005, 1e-7/1e-7, 100 its",
            "value": evals / (gpu_ms * 1e-3) / 1e6, "unit": "Mpts/s", "scaling": "pairs round-robin over ranks, no collective",
            "e2e": {"value": evals / wall / 1e6, "unit": "Mpts/s", "ms_per_pair": 1e3 * wall * world / npairs,
                    "step": "per pair: upload both clouds, voxel_down_sample, estimate_normals, registration_generalized_icp, result to host"},
            "iterations_total": int(iters), "gpu_launches": int(launches), "max_t_err_m": terr, "max_r_err_rad": rerr,
            "roofline": {"kernel": "k_gicp_linearize", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": measured_traffic("k_gicp_linearize_c4"), "peak_source": peak_src,
                         "note": "144 B per source point per linearisation (SURVEY.md 8d), per GPU; ~50 k-point clouds: launch/latency bound"},
            "cpu_baseline": {"value": ev / tr / 1e6, "unit": "Mpts/s", "cores": host_cores(), "kind": "port",
                             "sample": f"pair (1 -> 0): {it} iterations in {tr:.3f} s (registration only), {tt:.3f} s with downsampling + normals",
                             "e2e_value": ev / tt / 1e6}}
    # ---- C4, standard mode: the 4 pairs (every other lidar onto lidar 0) of multi_lidar_calibrator.py:202-219
    c4s = GB.run_c4(rank, world, mode="standard")
    wall_s = allmax(c4s["wall_s"]); gpu_ms_s = allmax(c4s["gpu_ms"]); evals_s = allsum(c4s["src_evals"]); npairs_s = allsum(c4s["pairs"])
    if rank == 0:
        out["gicp_c4_standard"] = {
            "workload": f"C4 Multi_LiCa GICP, standard mode: {int(npairs_s)} pairs onto lidar 0 (multi_lidar_calibrator.py:202-219), same parameters",
            "value": evals_s / max(gpu_ms_s, 1e-9) / 1e3, "unit": "Mpts/s", "scaling": "pairs round-robin over ranks (at most 4 ranks have work), no collective",
            "e2e": {"value": evals_s / wall_s / 1e6, "unit": "Mpts/s", "ms_per_pair": 1e3 * wall_s * min(world, max(int(npairs_s), 1)) / max(npairs_s, 1),
                    "step": "per pair: upload both clouds, voxel_down_sample, estimate_normals, registration_generalized_icp, result to host"}}
    # ---- C5: one registration, source sharded over the ranks, 30 doubles all-reduced per iteration
    comm = None
    if world > 1:
        ids = [gicp.Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = gicp.Communicator(ids[0], rank, world)
    c5 = GB.run_c5(args.c5_points, rank, world, comm)
    ms = allmax(c5["gpu_ms"])
    if rank == 0:
        ach = 144.0 * c5["points"] * c5["evaluations"] / world / (ms * 1e-3) / 1e9
        out["gicp_c5"] = {
            "workload": f"C5 map-to-map GICP: {c5['points']} source pts vs {c5['points']} target pts (city block tiled 4x4), "
                        f"{c5['iterations']} fixed iterations, max_corr 1.0", "value": c5["points"] * c5["evaluations"] / (ms * 1e-3) / 1e6,
            "unit": "Mpts/s", "n_gpus": world, "scaling": "strong", "ms_per_align": ms, "ms_per_evaluation": ms / c5["evaluations"],
            "parallelism": (f"target replicated, source sharded x{world} (Morton blocks), the 30 sums exchanged "
                            + ("inside k_gicp_linearize over NVLink peer memory" if c5.get("fused_exchange") else "by ncclAllReduce")
                            + "; upload and normals sharded + all-gathered") if world > 1 else "single GPU",
            "evaluation_ms": c5["evaluation_ms"],
            "e2e": {"value": c5["points"] * c5["evaluations"] / c5["e2e_s"] / 1e6, "unit": "Mpts/s", "s": c5["e2e_s"],
                    "step": "host clouds in: upload, estimate_normals (30-NN) on both, index builds, align, transformation out (this rank)",
                    "h2d_bytes": int(2 * 24 * c5["points"]), "d2h_bytes": 128 + 16},
            "gpu_launches": c5["launches"], "fitness": c5["fitness"], "inlier_rmse": c5["inlier_rmse"], "t_err_m": c5["t_err"],
            "r_err_rad": c5["r_err"], "target_cell_edge_m": c5["cell_edge"], "setup_s": c5["setup_s"],
            "roofline": {"kernel": "k_gicp_linearize", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": measured_traffic("k_gicp_linearize") if (world == 1 and c5["points"] == 50_000_000) else None,
                         "peak_source": peak_src, "bytes_per_launch": 144.0 * c5["points"] / world,
                         "note": "144 B per source point per linearisation (SURVEY.md 8d), per GPU, over the whole align (in the first evaluations, "
                                 "at the 0.3 m / 0.8 deg offset, most queries go through the coarse pass and have no target point within the "
                                 "radius: they are instruction-bound, see evaluation_ms)"}}
        if args.c5_cpu_points > 0:
            from oracle import pyoracle as O
            n = args.c5_cpu_points
            s_, t_, _ = GB.c5_clouds_torch(n, "cuda")
            cores = host_cores()
            _, sc = O.gicp_normals_covs(s_, 30, 0.005, cores)
            _, tc = O.gicp_normals_covs(t_, 30, 0.005, cores)
            go = O.GicpOracle(s_, sc, t_, tc, cores)
            t0 = time.perf_counter(); reps = 0
            while time.perf_counter() - t0 < 6.0:
                go.linearize(np.eye(4), 1.0); reps += 1
            el = time.perf_counter() - t0
            out["gicp_c5"]["cpu_baseline"] = {"value": n * reps / el / 1e6, "unit": "Mpts/s", "cores": cores, "kind": "port",
                                              "sample": f"{reps} linearisations of a {n}-point sample of the same scene in {el:.1f} s (kd-tree prebuilt)"}
    return out if rank == 0 else None


_RESULT_OUT = None


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--c5-points", type=int, default=50_000_000, help="points of the sharded map-to-map GICP (config C5)")
    ap.add_argument("--c5-cpu-points", type=int, default=500_000, help="sample size of the C5 cpu_baseline (0 = skip)")
    ap.add_argument("--skip-registration", action="store_true", help="C1 only (skip the NDT / GICP workloads)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: libraries that chat on file descriptor 1 (NCCL prints its version
    # there) are sent to stderr, and the result is written to a private copy of the original stdout
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
