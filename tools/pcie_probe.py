#!/usr/bin/env python
"""Host <-> device copy bandwidth with every rank copying at once (no kernels): what the host gives N concurrent GPUs.
Evidence for the end-to-end scaling numbers of bench.py (`e2e`): if the bare copies of the same sizes already run at the
per-GPU rate the pipeline sees, the limit is the host side (memory placement / socket link / PCIe switches), not the path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29566 tools/pcie_probe.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from cerberusnet_b200 import parallel
numa = parallel.bind_to_gpu_numa(local, local, world)   # same binding as bench.py
H2D, D2H = 15_572_992, 28_283_904                       # bytes per step of the bench's end-to-end pipeline
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory(); h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev); d_out = torch.empty(D2H, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=200):
    def step():
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"ms": round(dt * 1e3, 4), "h2d_GBps": round(H2D / dt / 1e9, 1) if kind != "d2h" else None,
            "d2h_GBps": round(D2H / dt / 1e9, 1) if kind != "h2d" else None}


res = {k: run(k) for k in ("h2d", "d2h", "both")}
res["numa"] = numa
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, res)
else:
    allr = [res]
if rank == 0:
    both = [r["both"] for r in allr]
    print(json.dumps({"n_gpus": world, "bytes_per_step": {"h2d": H2D, "d2h": D2H},
                      "concurrent_copies_per_rank": both,
                      "sum_GBps_both_directions": round(sum(b["h2d_GBps"] + b["d2h_GBps"] for b in both), 1),
                      "h2d_alone_per_rank_GBps": [r["h2d"]["h2d_GBps"] for r in allr],
                      "d2h_alone_per_rank_GBps": [r["d2h"]["d2h_GBps"] for r in allr], "numa_rank0": numa}))
if world > 1:
    dist.destroy_process_group()
