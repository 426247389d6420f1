"""A/B of the L2 eviction hints (and the evict_last budget) on the tcgen05 op at BASELINE config 2 (profiling aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_backward, cca_forward
lib = capi.load()
hint = lib.cca_b200__set_l2_hints; hint.argtypes = [ctypes.c_int, ctypes.c_double]; hint.restype = None
dev = torch.device("cuda:0")
B, Cq, C, H, W = 8, 64, 512, 97, 97
cl = torch.channels_last
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=10):
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
    out, lse = cca_forward(q, k, v, impl="tc")
    for on, mb in ((0, 0), (1, 0), (1, 45), (1, 70), (1, 90), (1, 110), (1, 140)):
        hint(on, float(mb))
        tf = timeit(lambda: cca_forward(q, k, v, impl="tc"))
        tb = timeit(lambda: cca_backward(do, q, k, v, out, lse, impl="tc"))
        def both():
            o, l = cca_forward(q, k, v, impl="tc"); cca_backward(do, q, k, v, o, l, impl="tc")
        tfb = timeit(both)
        print({"dtype": str(dt).split(".")[1], "hints": on, "keep_mb": mb, "fwd_ms": tf, "bwd_ms": tb, "fwd+bwd_ms": tfb}, flush=True)
