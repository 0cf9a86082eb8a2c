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
    json.dump({"T": r.transformation.tolist(), "iterations": r.iterations, "fitness": r.fitness, "rmse": r.inlier_rmse},
              open(os.path.join(work, f"rank{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
