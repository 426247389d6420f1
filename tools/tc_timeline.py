"""Dump the per-role clock64 timeline of CTA 0 of the tcgen05 forward kernel (profiling aid)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_forward
lib = capi.load()
dev = torch.device("cuda:0")
B, Cq, C, H, W = 8, 64, 512, 97, 97
q = torch.randn(B, Cq, H, W, device=dev).contiguous(memory_format=torch.channels_last)
k = torch.randn(B, Cq, H, W, device=dev).contiguous(memory_format=torch.channels_last)
v = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
for _ in range(3):
    cca_forward(q, k, v, impl="tc")
buf = torch.zeros(2 * 5 * 512, dtype=torch.int64, device=dev)
fn = lib.cca_b200__set_debug_buffer
fn.argtypes = [ctypes.c_void_p]; fn.restype = None
tp = lib.cca_b200__set_two_pass; tp.argtypes = [ctypes.c_int]; tp.restype = None
tp(int(sys.argv[1]) if len(sys.argv) > 1 else 1)        # 1: two launches (default), 2: static fused, 0: dynamic fused
for _ in range(2):
    cca_forward(q, k, v, impl="tc")
fn(buf.data_ptr())
cca_forward(q, k, v, impl="tc")
torch.cuda.synchronize()
fn(None)
import time
def timeit(tag):
    for _ in range(3): cca_forward(q, k, v, impl="tc")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): cca_forward(q, k, v, impl="tc")
    e1.record(); torch.cuda.synchronize()
    print(tag, "fwd ms", e0.elapsed_time(e1) / 20)
tp(0); timeit("one launch, dynamic :")
tp(2); timeit("one launch, static  :")
tp(1); timeit("two launches        :")
t = buf.cpu().view(2, 5, 512)
names = ["producer(slot free->issue)", "converter(full, op_empty, done)", "mma", "softmax", "epilogue(top, out_full, o_full, tmem_ld done, sts done, staged)"]
for ps, pname in enumerate(["FUSED / COLUMN pass", "ROW pass"]):
    vals = [int(x) for x in t[ps].flatten() if x > 0]
    if not vals:
        continue
    base = min(vals)
    print("=====", pname)
    for role in range(5):
        st = [int(x) - base for x in t[ps, role] if x > 0]
        print(f"--- {names[role]}: {len(st)} stamps")
        print(" ".join(str(x) for x in st))
