"""Run the op forward+backward at BASELINE config 2 a few times (target for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import cca_backward, cca_forward
dev = torch.device("cuda:0")
B, Cq, C, H, W = 8, 64, 512, 97, 97
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
cl = torch.channels_last
q = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
k = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
v = (torch.randn(B, C, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
do = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=cl)
for _ in range(iters):
    out, lse = cca_forward(q, k, v, impl="tc")
    dq, dk, dv = cca_backward(do, q, k, v, out, lse, impl="tc")
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(dv.float().abs().mean()))
