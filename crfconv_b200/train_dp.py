"""Data-parallel training step for PointConvResNet — the loop of the reference's ``Trainer.train_one_epoch``
(trainval.py:92-108) with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        -m crfconv_b200.train_dp --steps 20 --clouds-per-gpu 2 --points 40960

Each rank builds (or receives) its own slice of the batch — clouds are independent in every stage — runs the multiscale builder,
forward, cross-entropy and backward locally, and the ranks exchange ONE all-reduce of the flat gradient buffer (3.28 MB) before
the SGD step.  Optimiser, schedule and loss follow the reference (SGD + momentum + weight decay, ExponentialLR, class-weighted
cross-entropy with ignore_index).  Data here is synthetic (there is no dataset on the box); ``train_step`` is what a real loader
would call.
"""
from __future__ import annotations

import argparse
import os
import time
import types

import torch
import torch.distributed as dist

from . import losses
from .distributed import FlatGradients, init_from_env
from .multiscale import build_multiscale
from .point_conv_big import PointConvResNet


def make_batch(pos, feats, labels, generator=None):
    """pos [B,N,3], feats [B,N,C], labels [B,N] (CUDA) → the object ``PointConvResNet.forward`` reads (data.x, data.multiscale, data.y)."""
    return types.SimpleNamespace(x=feats, y=labels, multiscale=build_multiscale(pos, generator=generator))


def train_step(model, grads: FlatGradients, optimizer, data, class_weights=None, ignore_index=-1, world_size=1):
    """One optimisation step on this rank's shard; returns the local loss (a CUDA scalar, no host sync).

    The reference's loss is a class-weighted mean over the non-ignored points of the WHOLE batch (trainval.py:100-104).  With the
    batch sharded, each rank back-propagates the weighted SUM over its shard; its normaliser (Σ of the class weights of its
    non-ignored points) rides in the trailing slot of the flat gradient buffer, so the one all-reduce yields both the summed
    gradient and the global normaliser — uneven shards, ignore_index and class weights all reproduce the single-process gradient."""
    grads.zero()
    y_pred = model(data)
    y_target = data.y.reshape(-1) - 1                                    # trainval.py:100
    # one forward and one backward kernel (csrc/loss.cu); the normaliser Σ w[t] over the non-ignored points comes with the forward
    loss_sum, norm = losses.cross_entropy(y_pred, y_target, weight=class_weights, ignore_index=ignore_index, reduction="sum",
                                          return_normaliser=True)
    norm = norm.to(torch.float32)
    loss_sum.backward()
    if world_size > 1 and grads.extra.numel():
        grads.extra[0] = norm
        grads.all_reduce(average=False)                                  # the only collective of the step
        grads.flat[:grads.numel].div_(grads.extra[0].clamp(min=1e-12))
    else:
        if world_size > 1:
            grads.all_reduce(average=False)
            norm = norm * world_size                                     # equal shards assumed when no normaliser slot was reserved
        grads.flat[:grads.numel].div_(norm.clamp(min=1e-12))
    optimizer.step()
    return (loss_sum / norm.clamp(min=1e-12)).detach()


def synthetic_shard(clouds, points, n_classes, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    pos = torch.rand(clouds, points, 3, generator=g) * torch.tensor([8.0, 6.0, 3.0])
    rgb = torch.rand(clouds, points, 3, generator=g)
    labels = torch.randint(1, n_classes + 1, (clouds, points), generator=g)
    pos = pos.to(device)
    return pos, torch.cat([pos, rgb.to(device)], dim=-1), labels.to(device), g


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--clouds-per-gpu", type=int, default=2)
    ap.add_argument("--points", type=int, default=40960)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--momentum", type=float, default=0.98)
    ap.add_argument("--weight-decay", type=float, default=1e-4)
    ap.add_argument("--gamma", type=float, default=0.95)
    ap.add_argument("--epoch-steps", type=int, default=500, help="optimisation steps per epoch: ExponentialLR steps once per epoch (trainval.py)")
    ap.add_argument("--checkpoint", default="")
    args = ap.parse_args(argv)

    rank, local_rank, world = init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(0)                                                 # identical initial weights on every rank
    model = PointConvResNet(in_channels=6, n_classes=args.classes, use_crf=True, steps=1).to(dev).train()
    grads = FlatGradients(model, extra=1, direct=True)
    opt = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=args.gamma)
    pos, feats, labels, gen = synthetic_shard(args.clouds_per_gpu, args.points, args.classes, dev, seed=1000 + rank)
    t0 = time.time()
    for step in range(args.steps):
        data = make_batch(pos, feats, labels, generator=gen)
        loss = train_step(model, grads, opt, data, world_size=world)
        if step == 0:
            torch.cuda.synchronize(dev)
            t0 = time.time()                                             # first step pays for lazy initialisation
        if (step + 1) % args.epoch_steps == 0:
            sched.step()                                                 # once per epoch, like the reference's Trainer
    torch.cuda.synchronize(dev)
    dt = (time.time() - t0) / max(args.steps - 1, 1)
    if world > 1:
        w = torch.cat([p.detach().flatten()[:64] for p in model.parameters()][:4])
        ref = w.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(w, ref), "replicas diverged"                  # same gradients ⇒ bit-identical parameters
    if rank == 0:
        print(f"train_dp: world={world} clouds/gpu={args.clouds_per_gpu} points={args.points} loss={float(loss):.4f} "
              f"{dt * 1e3:.1f} ms/step  {world * args.clouds_per_gpu * args.points / dt / 1e6:.2f} M points/s")
        if args.checkpoint:
            model.save(args.checkpoint)                                  # Base.save (common.py:89-92): plain state_dict
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
