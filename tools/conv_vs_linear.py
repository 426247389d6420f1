"""Microbenchmark: the module's three 1x1 projections as cudnn convs vs F.linear on the channels-last view (fp32, no TF32)."""
import torch, torch.nn as nn, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, C, H, W = 8, 512, 97, 97
Cq = C // 8
convs = [nn.Conv2d(C, Cq, 1).to(dev), nn.Conv2d(C, Cq, 1).to(dev), nn.Conv2d(C, C, 1).to(dev)]
x = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)

def run_conv():
    outs = [c(x) for c in convs]
    sum(o.sum() for o in outs).backward()

def run_linear():
    xm = x.permute(0, 2, 3, 1).reshape(-1, C)
    outs = [F.linear(xm, c.weight.view(c.out_channels, C), c.bias) for c in convs]
    sum(o.sum() for o in outs).backward()

def run_linear_cat():
    xm = x.permute(0, 2, 3, 1).reshape(-1, C)
    w = torch.cat([c.weight.view(c.out_channels, C) for c in convs]); b = torch.cat([c.bias for c in convs])
    F.linear(xm, w, b).sum().backward()

for name, fn in (("conv2d x3", run_conv), ("linear x3", run_linear), ("linear cat", run_linear_cat)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:12s} fwd+bwd {e0.elapsed_time(e1)/10:.3f} ms")
