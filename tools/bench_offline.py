#!/usr/bin/env python3
"""Config C4: offline-phase stress, M pcore MCs x D dims, pairwise eps-neighbourhood row-sharded over the
ranks (launch with torchrun for G > 1).  Prints one JSON line on rank 0.

    python tools/bench_offline.py --M 100000 --D 40
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_offline.py ...
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chronoclust_b200 import _lib  # noqa: E402
from chronoclust_b200.offline_sharded import CudaStages, sharded_offline  # noqa: E402
from chronoclust_b200.synth import gen_offline_stress  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=100000)
    ap.add_argument("--D", type=int, default=40)
    ap.add_argument("--E", type=float, default=0.3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cen, w, core = gen_offline_stress(a.M, a.D)
    tc = torch.from_numpy(cen).cuda()
    tcore = torch.from_numpy(core.astype(np.uint8)).cuda()
    st = CudaStages(local, dnrm2_ptr=_lib.scipy_dnrm2_pointer())
    times = []
    for rep in range(a.reps + 1):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lab, order, cl_off, ncl, info = sharded_offline(st, tc, tcore, a.M, a.D, 4.0, a.D, 0.05, a.E, a.E ** 2, dist=dist)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep:
            times.append(float(t[0]))
    stage_s = {}
    sharded_offline(st, tc, tcore, a.M, a.D, 4.0, a.D, 0.05, a.E, a.E ** 2, dist=dist, timers=stage_s)  # one extra pass, per-stage times
    if rank == 0:
        ms = float(np.mean(times))
        pairs = float(a.M) ** 2
        print(json.dumps({"metric": "offline ordered MC pairs/s (eps-neighbourhood + subspace + weighted + clusters)",
                          "value": pairs / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "ms": ms, "M": a.M, "D": a.D,
                          "euclid_gflops": 3.0 * a.D * pairs / (ms * 1e-3) / 1e9, "clusters": int(ncl),
                          "clustered_mcs": int((lab >= 0).sum()), "info": info, "scaling": "strong",
                          "stage_ms_rank0": {k: round(v * 1e3, 2) for k, v in stage_s.items()}}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
