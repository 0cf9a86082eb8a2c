"""One rank of the 2-GPU sharded GICP test (launched by torchrun from tests/test_gpu_gicp.py)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_sensor_slam_tookit_b200 import capi, gicp  # noqa: E402


def main():
    work = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    capi.check(capi.lib().b2_set_device(local))
    dist.init_process_group("gloo")
    ids = [gicp.Communicator.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = gicp.Communicator(ids[0], rank, world)
    assert np.array_equal(comm.allreduce(np.arange(4.0) + rank), world * np.arange(4.0) + sum(range(world)))
    d = np.load(os.path.join(work, "in.npz"))
    s = gicp.PointCloud(d["sp"]); s.normals = d["sn"]
    t = gicp.PointCloud(d["tp"]); t.normals = d["tn"]
    g = gicp.GeneralizedICP(1.0, 0.005)
    g.setInputTarget(t); g.setInputSource(s)
    g.setShard(comm)
    r = g.align(d["init"])
    # the sharded set-up of config C5: slice upload + all-gather, sharded normals + all-gather, source grid over this rank's rows
    s2 = gicp.PointCloud.from_host_sharded(d["sp"], comm); t2 = gicp.PointCloud.from_host_sharded(d["tp"], comm)
    pts_equal = bool(np.array_equal(s2.points, d["sp"]) and np.array_equal(t2.points, d["tp"]))
    s2.estimate_normals_sharded(comm); t2.estimate_normals_sharded(comm)
    s1 = gicp.PointCloud(d["sp"]); s1.estimate_normals()
    nrm_equal = bool(np.array_equal(s2.normals, s1.normals))
    g2 = gicp.GeneralizedICP(1.0, 0.005)
    g2.setInputTarget(t2)
    b, e = gicp.row_slice(len(d["sp"]), rank, world)
    g2.setInputSourceSlice(s2, b, e)
    g2.setShard(comm)
    r2 = g2.align(d["init"], want_correspondences=False)
    # the same with the exchange fused into the kernel (peer stores over NVLink instead of ncclAllReduce + an epilogue launch)
    hs = [None] * world
    dist.all_gather_object(hs, g2.peerHandle())
    fused = g2.setPeers(rank, world, hs)
    r3 = g2.align(d["init"], want_correspondences=False) if fused else r2
    r4 = g2.align(d["init"], want_correspondences=False) if fused else r2        # a second align on the same handle (sequence numbers go on)
    # Morton-block shards (the deal the C5 bench uses)
    g5 = gicp.GeneralizedICP(1.0, 0.005)
    g5.setInputTarget(t2); g5.setInputSourceBlocks(s2, rank, world); g5.setShard(comm)
    r5 = g5.align(d["init"], want_correspondences=False)
    # ... and the one-call exchange set-up (handles all-gathered over the communicator) on the block shards
    fused6 = g5.setupExchange()
    r6 = g5.align(d["init"], want_correspondences=False) if fused6 else r5
    json.dump({"T": r.transformation.tolist(), "iterations": r.iterations, "fitness": r.fitness, "rmse": r.inlier_rmse,
               "pts_equal": pts_equal, "nrm_equal": nrm_equal, "T2": r2.transformation.tolist(), "iterations2": r2.iterations,
               "fitness2": r2.fitness, "slice": [b, e],
               "T5": r5.transformation.tolist(), "iterations5": r5.iterations, "fitness5": r5.fitness, "fused": bool(fused), "T3": r3.transformation.tolist(), "iterations3": r3.iterations, "T4": r4.transformation.tolist(),
               "fused6": bool(fused6), "T6": r6.transformation.tolist(), "iterations6": r6.iterations},
              open(os.path.join(work, f"rank{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
