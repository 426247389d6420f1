"""Synthetic train-step harness for BASELINE configs[2] (ResNet101+RCCA forward on 769x769, 1 GPU) and configs[3] (DDP train
step on synthetic Cityscapes-shaped data, global batch 8, R=2) -- SURVEY.md 8(f) N2.

Mirrors the reference's loop on current torch (train.py:199-239, engine.py:49-99): one process per GPU, per-GPU batch =
global batch / world (engine.py:88), stock DistributedDataParallel, BatchNorm -> SyncBatchNorm when world > 1 (the reference's
InPlaceABNSync), SGD(momentum 0.9, weight decay 5e-4) with the poly schedule, loss = CriterionDSN (loss/criterion.py:10-34),
`optimizer.zero_grad(); loss = model(images, labels); all_reduce(loss); loss.backward(); optimizer.step()`.
Images randn[b,3,crop,crop]; labels randint(0,19) with a 255 "ignore" frame, as SURVEY.md 8(d) prescribes.

  python harness/train_synth.py [--steps 5 --warmup 2 --crop 769 --batch 8 --recurrence 2]      (or under torchrun)
prints one JSON line; `run()` is what bench.py imports for its `ccnet` key."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _events_ms(fn, steps, warmup, barrier=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(local_rank: int, world: int, steps: int = 5, warmup: int = 2, crop: int = 769, global_batch: int = 8,
        recurrence: int = 2, num_classes: int = 19, lr: float = 1e-2, forward_only_too: bool = True, allow_tf32: bool = False):
    from harness.ccnet_model import CCNet, dsn_loss
    from ccnet_b200 import RCCA
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.benchmark = True
    per_gpu = max(1, global_batch // world)                       # engine.py:88
    torch.manual_seed(7 + local_rank)
    net = CCNet(num_classes=num_classes, recurrence=recurrence).to(dev).to(memory_format=torch.channels_last)
    with torch.no_grad():
        net.head.cca.gamma.fill_(0.5)                             # gamma = 0 at init would make the operator's output vanish
    if world > 1:
        # (torch.nn.SyncBatchNorm copies a mask to the host in every layer's forward: harness/sync_bn.py; HARNESS_SYNCBN=torch selects it)
        if os.environ.get("HARNESS_SYNCBN", "") == "torch":
            net = nn.SyncBatchNorm.convert_sync_batchnorm(net)
        else:
            from harness.sync_bn import convert_sync_batchnorm
            net = convert_sync_batchnorm(net)
        model = nn.parallel.DistributedDataParallel(net, device_ids=[local_rank])
    else:
        model = net
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=5e-4)     # train.py:161-163
    images = torch.randn(per_gpu, 3, crop, crop, device=dev).contiguous(memory_format=torch.channels_last)
    labels = torch.randint(0, num_classes, (per_gpu, crop, crop), device=dev)
    labels[:, :8, :] = 255; labels[:, -8:, :] = 255; labels[:, :, :8] = 255; labels[:, :, -8:] = 255
    barrier = (lambda: dist.barrier()) if world > 1 else None
    it = {"n": 0}
    total_iters = 60000

    def train_step():
        opt.zero_grad(set_to_none=True)
        for g in opt.param_groups:                                # train.py adjust_learning_rate: poly, power 0.9
            g["lr"] = lr * (1 - it["n"] / total_iters) ** 0.9
        loss = dsn_loss(model(images), labels)
        if world > 1:
            dist.all_reduce(loss.detach().clone())               # engine.all_reduce_tensor(loss)
        loss.backward()
        opt.step()
        it["n"] += 1

    out = {"crop": crop, "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "recurrence": recurrence,
           "precision": "fp32" + (" (tf32 convs)" if allow_tf32 else ""), "layout": "channels_last",
           "model": "ResNet101 + RCCA (harness/ccnet_model.py, state-dict compatible with networks/ccnet.py)"}
    ms_train = _events_ms(train_step, steps, warmup, barrier)
    if world > 1:
        t = torch.tensor([ms_train], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_train = float(t.item())
    out["train_ms_per_step"] = ms_train
    out["train_images_per_s"] = per_gpu * world / (ms_train * 1e-3)
    # share of the criss-cross attention modules (projections + operator + residual, R times, fwd+bwd) in that step: the head
    # sees [per_gpu, 512, crop/8 + 1, crop/8 + 1]
    hw = (crop - 1) // 8 + 1
    rc = RCCA(512, recurrence=recurrence).to(dev)
    rc.cca.load_state_dict(net.head.cca.state_dict())      # `net` is the bare network, `model` its DDP wrapper
    xh = torch.randn(per_gpu, 512, hw, hw, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gh = torch.randn_like(xh)

    def cca_step():
        y = rc(xh)
        y.backward(gh)
        xh.grad = None
        rc.zero_grad(set_to_none=True)

    ms_cca = _events_ms(cca_step, max(steps, 5), 2)
    out["cca_modules_ms"] = ms_cca
    out["cca_share_of_train_step"] = ms_cca / ms_train
    out["head_feature_map"] = [per_gpu, 512, hw, hw]
    if forward_only_too and local_rank == 0:
        net_eval = net
        x1 = images[:1]

        def fwd():
            with torch.no_grad():
                net_eval(x1)

        net_eval.eval()
        out["forward_ms_batch1"] = _events_ms(fwd, max(steps, 5), 2)      # BASELINE configs[2]
        net_eval.train()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--crop", type=int, default=769)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--recurrence", type=int, default=2)
    ap.add_argument("--tf32", action="store_true")
    a = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r = run(local_rank, world, a.steps, a.warmup, a.crop, a.batch, a.recurrence, allow_tf32=a.tf32)
    if rank == 0:
        print(json.dumps({"harness": "ccnet_train_synth", "n_gpus": world, **r}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
