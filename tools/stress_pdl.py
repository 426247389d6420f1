"""Stress of the default launch mode (programmatic dependent launch): N forward+backward calls, each compared bit-for-bit
with the result of plain serial launches (profiling / robustness aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_backward, cca_forward
lib = capi.load()
pdl = lib.cca_b200__set_pdl; pdl.argtypes = [ctypes.c_int]; pdl.restype = None
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cl = torch.channels_last
bad = 0
for dt in (torch.float32, torch.bfloat16):
    for (B, Cq, C, H, W) in ((8, 64, 512, 97, 97), (3, 32, 256, 20, 97), (1, 16, 64, 9, 5)):
        q = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
        k = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
        v = (torch.randn(B, C, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
        do = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=cl)
        pdl(0)
        ro, rl = cca_forward(q, k, v, impl="tc")
        rg = cca_backward(do, q, k, v, ro, rl, impl="tc")
        pdl(1)
        fails = 0
        for it in range(N):
            o, l = cca_forward(q, k, v, impl="tc")
            g = cca_backward(do, q, k, v, o, l, impl="tc")
            if it % 7 == 0:                                  # some unrelated work in between, as in a real network
                (q * 1.0001).sum().item()
            ok = torch.equal(o, ro) and torch.equal(l, rl) and all(torch.equal(a, b) for a, b in zip(g, rg))
            fails += 0 if ok else 1
        bad += fails
        print({"dtype": str(dt).split(".")[1], "shape": (B, Cq, C, H, W), "calls": N, "mismatches": fails}, flush=True)
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
