"""Data-parallel training of the configs[4] skeleton through the package's kernels (projection forward, training forward of the
memory op, `gdr_bwd_kernel`), the way upstream trains (reference website/src/pages/[lang]/reprod/index.astro:238-252: 2-GPU DDP,
lr 1e-4): one process per GPU, NCCL gradient all-reduce by torch DistributedDataParallel, synthetic clips and masks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 examples/train_ddp.py --steps 20

Prints one JSON line from rank 0: first / last loss, steps per second, and whether the parameters of all ranks are identical.
The stand-in layers around the memory are random-init PyTorch: this shows that the op trains under DDP,
not that the model reproduces the paper."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--clips", type=int, default=2, help="clips per GPU and step")
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--lr", type=float, default=1e-4)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gdkvm_b200.model import GDKVMSkeleton
    import gdkvm_b200
    torch.manual_seed(0)                                    # same initial weights on every rank
    model = GDKVMSkeleton().to(dev)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=args.lr)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)                  # different data per rank
    # a fixed synthetic task: the mask is a disc whose radius follows the clip's mean intensity -- learnable, so the loss falls
    clips = torch.rand(args.clips, args.frames, 1, 112, 112, generator=gen, device=dev)
    yy, xx = torch.meshgrid(torch.arange(112, device=dev), torch.arange(112, device=dev), indexing="ij")
    rad = 20 + 30 * clips.mean(dim=(2, 3, 4))                                  # [clips, frames]
    masks = (((yy - 56) ** 2 + (xx - 56) ** 2)[None, None] < rad[..., None, None] ** 2).float()[:, :, None]
    losses = []
    n0 = gdkvm_b200.launch_count()
    torch.cuda.synchronize()
    t0 = time.time()
    for step in range(args.steps):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, _ = net(clips)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float(), masks)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    torch.cuda.synchronize()
    dt = time.time() - t0
    # identical parameters on every rank after the run = the gradients were all-reduced
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
    same = True
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([float(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({"world_size": world, "steps": args.steps, "loss_first": losses[0], "loss_last": losses[-1],
                          "steps_per_s": args.steps / dt, "frames_per_s": args.steps * args.clips * args.frames * world / dt,
                          "params_identical_across_ranks": same, "kernel_launches_rank0": gdkvm_b200.launch_count() - n0,
                          "grad_finite": bool(torch.isfinite(flat).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
