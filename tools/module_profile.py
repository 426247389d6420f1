"""Kernel-level breakdown of the user-facing module step (R=2 fwd+bwd at BASELINE config 2) with torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from ccnet_b200 import RCCA
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, C, H, W, R = 8, 512, 97, 97, 2
m = RCCA(C, recurrence=R).to(dev)
with torch.no_grad():
    m.cca.gamma.fill_(1.0)
x = torch.randn(B, C, H, W, device=dev, requires_grad=True)
g = torch.randn(B, C, H, W, device=dev)
def step():
    y = m(x)
    (y * g).sum().backward()
    x.grad = None
    m.zero_grad(set_to_none=True)
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5.0, e.count // 5) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows if not r[0].startswith(("aten::", "autograd::", "_FusedCCAStep", "ProfilerStep")))
print("device time per step (us), kernels only: %.0f" % tot)
for k, t, n in rows[:40]:
    print("%10.1f us  x%-3d %s" % (t, n, k[:110]))
