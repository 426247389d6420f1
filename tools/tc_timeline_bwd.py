"""Per-role clock64 timeline of CTA 0 of the tcgen05 backward kernel (profiling aid)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_forward, cca_backward
lib = capi.load()
dev = torch.device("cuda:0")
B, Cq, C, H, W = 8, 64, 512, 97, 97
cl = torch.channels_last
q = torch.randn(B, Cq, H, W, device=dev).contiguous(memory_format=cl)
k = torch.randn(B, Cq, H, W, device=dev).contiguous(memory_format=cl)
v = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=cl)
do = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=cl)
out, lse = cca_forward(q, k, v, impl="tc")
for _ in range(3):
    cca_backward(do, q, k, v, out, lse, impl="tc")
buf = torch.zeros(2 * 5 * 512, dtype=torch.int64, device=dev)
fn = lib.cca_b200__set_bwd_debug_buffer
fn.argtypes = [ctypes.c_void_p]; fn.restype = None
fn(buf.data_ptr())
cca_backward(do, q, k, v, out, lse, impl="tc")
torch.cuda.synchronize()
fn(None)
t = buf.cpu().view(2, 5, 512)
names = ["producer(slot free)", "converter(full, op_empty, done)", "mma(line top, S issued, P_FULL got, [chunk done]x8, DS_FULL got, line end)",
         "P/dS(top, S_FULL got, P written, DP_FULL got, dS written)", "epilogue(top, ready, stored)"]
for ps, pname in enumerate(["COLUMN pass", "ROW pass"]):
    vals = [int(x) for x in t[ps].flatten() if x > 0]
    if not vals: continue
    base = min(vals)
    print("=====", pname)
    for role in range(5):
        st = [int(x) - base for x in t[ps, role] if x > 0]
        print(f"--- {names[role]}: {len(st)} stamps")
        print(" ".join(str(x) for x in st[:66]))
