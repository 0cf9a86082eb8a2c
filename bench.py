#!/usr/bin/env python
"""bench.py — scan-to-map LM throughput on the C1 workload of BASELINE.json (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

One JSON line on stdout (rank 0). A "step" is one scan2MapOptimization call: the LM loop of
liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:1282-1310 on one VLP-16 scan against the 100 000-point map.
  value      LM iterations/s with map and scan already resident in HBM (device time, CUDA events on the library's
             stream, L2 flushed between steps)
  e2e        the same through the C ABI with HOST buffers: set_map + set_scan + solve per step, wall clock
  roofline   k_s2m_iteration: algorithmic bytes (72 B + 16 B per candidate in the 27-cell block, per feature) over
             the measured launch time, against MEASURED_PEAKS.json's HBM copy bandwidth
  batched    B scans (pose hypotheses) against the one map in a single launch per iteration
  cpu_baseline  the CPU oracle (restatement of the reference path) on the box's host cores
--impl reference times that CPU path alone (PCL/OpenCV are not installable here: the oracle is the reference arm).
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
    cores = os.cpu_count() or 1
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
    iters, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        iters += step()
    el = time.perf_counter() - t0
    val = iters / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(c1), "step": "kd-tree build (both maps) + scan2MapOptimization loop",
                       "iters_per_step": iters / args.steps},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full steps (index build + LM loop) of the same workload, OpenMP over features "
                                       f"as mapOptmization.cpp:978,1070, {cores} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


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
    # ---- value: resident inputs, device time of each step (events on the library's stream)
    barrier()
    l0 = L.b2_kernel_launch_count()
    dev_ms, wall_ms, iters = 0.0, 0.0, 0
    for _ in range(K):
        flush_l2()
        t0 = time.perf_counter()
        r = solve()
        wall_ms += (time.perf_counter() - t0) * 1e3
        dev_ms += g.lastGpuMs()[0]
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
    # ---- batched: B pose hypotheses of the scan against the resident map, one launch per iteration
    batched = None
    B = args.batch
    if B > 1:
        rng = np.random.default_rng(7)
        poses = np.tile(c1["pose_truth"], (B, 1)).astype(np.float32)
        poses[:, 3:] += rng.uniform(-0.15, 0.15, (B, 3)).astype(np.float32)
        poses[:, :3] += np.deg2rad(rng.uniform(-1.0, 1.0, (B, 3))).astype(np.float32)
        gb = ScanToMapOptimizer(max_batch=B)
        gb.setInputMap(mc, ms)
        gb.setInputScanBatch([sc] * B, [ss] * B)
        rb = gb.scan2MapOptimizationBatch(poses)
        b_max = int(rb["iters"].max())
        for _ in range(3):
            flush_l2(); gb.scan2MapOptimizationBatch(poses, b_max)
        Kb = max(3, min(K, 20))
        b_ms, b_iters = 0.0, 0
        for _ in range(Kb):
            flush_l2()
            rb = gb.scan2MapOptimizationBatch(poses, b_max)
            b_ms += gb.lastGpuMs()[0]
            b_iters += int(rb["iters"].sum())
        batched = {"scans": B, "steps": Kb, "value": b_iters / (b_ms * 1e-3), "unit": UNIT, "ms_per_step": b_ms / Kb,
                   "iters_per_step": b_iters / Kb, "converged": int(rb["converged"].sum()), "max_iters_in_batch": b_max}
        del gb
    barrier()
    clocks = sampler.stop() if sampler else None

    # ---- reductions over ranks: units summed, time = max
    t = torch.tensor([dev_ms, e2e_s, act_ms, wall_ms], dtype=torch.float64, device="cuda")
    u = torch.tensor([iters, e_iters], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dev_ms_m, e2e_s_m, act_ms_m, wall_ms_m = [float(v) for v in t.tolist()]
    iters_all, e_iters_all = [float(v) for v in u.tolist()]
    if rank == 0:
        peak, peak_src = peaks()
        abytes, cand, compulsory = algorithmic_bytes_per_iteration(c1, guess)
        launch_ms = act_ms_m / (K * iters_needed)                 # span / active launches (includes launch gaps + prepare)
        achieved = abytes / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": iters_all / (dev_ms_m * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_m / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(c1), "step": "scan2MapOptimization loop (<=30 LM iterations, stops on convergence)",
                       "iters_per_step": iters / K, "l2": "flushed between steps (256 MiB device write)",
                       "parallelism": f"replicas x{world} (C1 does not shard; no data-path collective)",
                       "timing": "device ms per step = CUDA events on the library stream around each solve",
                       "wall_ms_per_step": wall_ms_m / K},
            "e2e": {"value": e_iters_all / e2e_s_m, "unit": UNIT,
                    "h2d_bytes_per_step": int(mc.nbytes + ms.nbytes + sc.nbytes + ss.nbytes + 592 + 16),
                    "d2h_bytes_per_step": int(592 + 4 + 2 * 24), "ms_per_step": 1e3 * e2e_s_m / K,
                    "step": "b2_s2m_set_map + b2_s2m_set_scan + b2_s2m_solve with pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_s2m_iteration", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_per_launch": abytes, "compulsory_bytes_per_launch": compulsory,
                         "candidates_per_feature": cand, "launch_ms": launch_ms,
                         "note": "span of a solve capped at the iterations it needs / launches; working set (1.7 MB) is L2-resident "
                                 "after first touch, so this is L2+DRAM bytes over time, see DESIGN.md"},
            "clocks": clocks,
        }
        if batched:
            bb = abytes * batched["iters_per_step"] / (batched["ms_per_step"] * 1e-3) / 1e9
            batched["roofline"] = {"achieved": bb, "peak": peak, "unit": "GB/s", "frac": bb / peak,
                                   "note": "algorithmic bytes of all active (scan, iteration) pairs / device span of the step"}
            line["batched"] = batched
        # CPU baseline on this box's host cores, bounded sample
        cores = os.cpu_count() or 1
        v_solve, st1, el1, it1 = cpu_solve_rate(c1, cores, 6.0, False)
        v_full, st2, el2, _ = cpu_solve_rate(c1, cores, 6.0, True)
        v_4, _, _, _ = cpu_solve_rate(c1, min(4, cores), 3.0, False)
        line["cpu_baseline"] = {"value": v_solve, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{st1} solves of the same scan/map in {el1:.1f} s (kd-trees prebuilt), {it1} iterations each",
                                "with_index_build": v_full, "threads4": v_4,
                                "note": "CPU restatement of the reference path (oracle/), OpenMP over features as the reference"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
