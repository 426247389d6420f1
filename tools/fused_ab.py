"""A/B of the forward launch modes (two launches / static fused / dynamic fused) x L2 hints at BASELINE config 2."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_forward
lib = capi.load()
hint = lib.cca_b200__set_l2_hints; hint.argtypes = [ctypes.c_int, ctypes.c_double]; hint.restype = None
tp = lib.cca_b200__set_two_pass; tp.argtypes = [ctypes.c_int]; tp.restype = None
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
only = sys.argv[1:]          # e.g. "2 1" = static fused with hints, one call (for ncu)
for dt in (torch.float32, torch.bfloat16):
    q = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
    k = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
    v = (torch.randn(B, C, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
    if only:
        tp(int(only[0])); hint(int(only[1]), 80.0)
        for _ in range(3): cca_forward(q, k, v, impl="tc")
        torch.cuda.synchronize()
        break
    tp(1); hint(0, 80.0)
    ref = cca_forward(q, k, v, impl="tc")
    for mode, name in ((1, "two launches"), (2, "static fused"), (0, "dynamic fused")):
        for on in (0, 1):
            tp(mode); hint(on, 80.0)
            o = cca_forward(q, k, v, impl="tc")
            same = bool(torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]))
            print({"dtype": str(dt).split(".")[1], "mode": name, "hints": on, "fwd_ms": timeit(lambda: cca_forward(q, k, v, impl="tc")), "bit_identical": same}, flush=True)
tp(1); hint(1, 80.0)
