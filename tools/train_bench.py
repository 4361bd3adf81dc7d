#!/usr/bin/env python
"""Training-step harness for BASELINE.json configs[3]: flow network (pyramid encoder + PWC-style
decoder on the fused op, forward AND backward flow) + photometric loss + Adam, batch 8 per GPU of
synthetic 1024x512 image pairs, DistributedDataParallel over NCCL (one process per GPU).

    python tools/train_bench.py [--steps K] [--warmup W] [--batch 8]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...

Prints one JSON line (rank 0): frames/s of the whole job (image pairs per second), device-timed,
max over ranks.  The only collective is DDP's gradient all-reduce; the op itself shards over the
batch with no exchange.
"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import cerberusnet_b200 as cb
from cerberusnet_b200.decoder import FlowNetLite, photometric_loss
from cerberusnet_b200.parallel import max_over_ranks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = FlowNetLite().to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), weight_decay=1e-6)
    g = torch.Generator().manual_seed(1234 + rank)
    host = [(torch.rand(a.batch, 3, a.height, a.width, generator=g).pin_memory(),
             torch.rand(a.batch, 3, a.height, a.width, generator=g).pin_memory()) for _ in range(2)]
    n0 = cb.lib().cerb_launch_count()

    def step(i):
        h1, h2 = host[i % 2]
        img1, img2 = h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)   # H2D inside the step
        out = net(img1, img2, consistency=True)
        loss = photometric_loss(img1, img2, out["flow"]) + photometric_loss(img2, img1, out["flow_b"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for i in range(max(a.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        loss = step(i)
    last = float(loss.item())                                                         # D2H of the step result
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier()
    ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms), dev) / a.steps
    launches = cb.lib().cerb_launch_count() - n0
    if rank == 0:
        print(json.dumps({
            "metric": "flow-network training frames/s (image pairs/s)", "value": world * a.batch / (ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "flow_train_step_1024x512", "batch_per_gpu": a.batch, "params": n_params,
                       "model": "PyramidEncoder + FlowDecoder (fused warp+corr+LeakyReLU fwd/bwd, both flow directions)",
                       "optimizer": "Adam", "parallelism": f"ddp{world}", "h2d_bytes_per_step": 2 * a.batch * 3 * a.height * a.width * 4},
            "last_loss": last, "costvolume_launches": int(launches)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
