"""Batch sweep of the tcgen05 op at C=512, 97x97: small batches stay L2-resident when the calls run back to back, which
separates the per-CTA pipeline rate from the HBM-bound rate of the BASELINE batch (profiling aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_backward, cca_forward
lib = capi.load()
tp = lib.cca_b200__set_two_pass; tp.argtypes = [ctypes.c_int]; tp.restype = None
dev = torch.device("cuda:0")
Cq, C, H, W = 64, 512, 97, 97
cl = torch.channels_last
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, fl, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if not fl:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
for dt in (torch.float32, torch.bfloat16):
    for B in (3, 6, 8, 12):
        q = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
        k = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
        v = (torch.randn(B, C, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
        do = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=cl)
        out, lse = cca_forward(q, k, v, impl="tc")
        row = {"dtype": str(dt).split(".")[1], "B": B, "items_per_cta": B * 97 / 148}
        for name, mode in (("two", 1), ("static", 2)):
            tp(mode)
            row["fwd_%s_warm" % name] = round(timeit(lambda: cca_forward(q, k, v, impl="tc"), False), 4)
            row["fwd_%s_cold" % name] = round(timeit(lambda: cca_forward(q, k, v, impl="tc"), True), 4)
        tp(1)
        row["bwd_warm"] = round(timeit(lambda: cca_backward(do, q, k, v, out, lse, impl="tc"), False), 4)
        row["bwd_cold"] = round(timeit(lambda: cca_backward(do, q, k, v, out, lse, impl="tc"), True), 4)
        print(row, flush=True)
