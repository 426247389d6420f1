"""A/B of programmatic dependent launch between the passes of the tcgen05 op: 0 off, 1 dependent launch, 2 dependent launch
with the second pass overlapping the tail of the first (per-sample completion counters).  Results must be bit-identical."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_backward, cca_forward
lib = capi.load()
pdl = lib.cca_b200__set_pdl; pdl.argtypes = [ctypes.c_int]; pdl.restype = None
dev = torch.device("cuda:0")
B, Cq, C, H, W = 8, 64, 512, 97, 97
cl = torch.channels_last
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=12):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2], 4)
for dt in (torch.float32, torch.bfloat16):
    q = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
    k = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
    v = (torch.randn(B, C, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
    do = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=cl)
    pdl(0)
    ref_o, ref_l = cca_forward(q, k, v, impl="tc")
    ref_g = cca_backward(do, q, k, v, ref_o, ref_l, impl="tc")
    for on in (0, 1, 2, 0, 1, 2):
        pdl(on)
        same = True
        for _ in range(5 if on < 2 else 40):
            o, l = cca_forward(q, k, v, impl="tc")
            g = cca_backward(do, q, k, v, o, l, impl="tc")
            same &= bool(torch.equal(o, ref_o) and torch.equal(l, ref_l) and all(torch.equal(a, b) for a, b in zip(g, ref_g)))
        tf = timeit(lambda: cca_forward(q, k, v, impl="tc"))
        tb = timeit(lambda: cca_backward(do, q, k, v, ref_o, ref_l, impl="tc"))
        print({"dtype": str(dt).split(".")[1], "pdl": on, "fwd_ms": tf, "bwd_ms": tb, "bit_identical_to_serial": same}, flush=True)
pdl(1)
