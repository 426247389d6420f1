"""Two-GPU check of harness/sync_bn.py (torchrun --nproc-per-node 2): (1) one layer against torch.nn.SyncBatchNorm -- output,
input / weight / bias gradients, running statistics; (2) the synthetic ResNet101+RCCA train step with either implementation."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dev = torch.device("cuda", local_rank)
from harness.sync_bn import convert_sync_batchnorm

torch.manual_seed(100 + rank)
res = {}
for fmt in (torch.channels_last, torch.contiguous_format):
    a, b = nn.BatchNorm2d(64).to(dev), nn.BatchNorm2d(64).to(dev)
    with torch.no_grad():
        w, bb = torch.randn(64, device=dev), torch.randn(64, device=dev)
        dist.broadcast(w, 0); dist.broadcast(bb, 0)
        for m in (a, b):
            m.weight.copy_(w); m.bias.copy_(bb)
    ref = nn.SyncBatchNorm.convert_sync_batchnorm(a)
    mine = convert_sync_batchnorm(b)
    x = (torch.randn(3 + rank, 64, 13, 17, device=dev) * 2 + 0.5).contiguous(memory_format=fmt)     # different counts per rank
    g = torch.randn_like(x)
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, ym = ref(xr), mine(xm)
    yr.backward(g); ym.backward(g)
    errs = {"y": (yr - ym).abs().max().item(), "dx": (xr.grad - xm.grad).abs().max().item(),
            "dw": (ref.weight.grad - mine.weight.grad).abs().max().item(), "db": (ref.bias.grad - mine.bias.grad).abs().max().item(),
            "rm": (ref.running_mean - mine.running_mean).abs().max().item(), "rv": (ref.running_var - mine.running_var).abs().max().item()}
    res["layer_" + ("nhwc" if fmt == torch.channels_last else "nchw")] = errs
    assert max(errs.values()) < 1e-5, errs

from harness.train_synth import run
for impl in ("nosync", "torch"):
    os.environ["HARNESS_SYNCBN"] = "torch" if impl == "torch" else ""
    r = run(local_rank, world, steps=3, warmup=2, allow_tf32=True, forward_only_too=False)
    res["train_" + impl] = {k: r[k] for k in ("train_ms_per_step", "train_images_per_s", "per_gpu_batch", "cca_modules_ms")}
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
