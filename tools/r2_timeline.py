"""Per-role clock64 stamps of CTA 0 of the forward values kernel and of the backward kernel (timeline build:
CCA_B200_LIBDIR=ccnet_b200/lib_tl CCA_B200_TIMELINE=1 CCA_B200_DEBUG_BUILD=1 python -m ccnet_b200.build, then run with
CCA_B200_LIB=ccnet_b200/lib_tl/libcca_b200.so).  Dumps the raw stamps as JSON for tools/r2_timeline_parse.py."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ccnet_b200 import capi, cca_backward, cca_forward
lib = capi.load()
dev = torch.device("cuda:0")
dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
B, Cq, C, H, W = 8, 64, 512, 97, 97
cl = torch.channels_last
q = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
k = (torch.randn(B, Cq, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
v = (torch.randn(B, C, H, W, device=dev) * 0.58).to(dt).contiguous(memory_format=cl)
do = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=cl)
for _ in range(3):
    out, lse = cca_forward(q, k, v, impl="tc")
    cca_backward(do, q, k, v, out, lse, impl="tc")
res = {}
for name, hook in (("stats", "cca_b200__set_stats_debug_buffer"), ("fwd", "cca_b200__set_debug_buffer"), ("bwd", "cca_b200__set_bwd_debug_buffer")):
    fn = getattr(lib, hook)
    fn.argtypes = [ctypes.c_void_p]; fn.restype = None
    buf = torch.zeros(5 * 512, dtype=torch.int64, device=dev)
    fn(buf.data_ptr())
    if name in ("fwd", "stats"):
        cca_forward(q, k, v, impl="tc")
    else:
        cca_backward(do, q, k, v, out, lse, impl="tc")
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().view(5, 512)
    vals = [int(x) for x in t.flatten() if x > 0]
    base = min(vals) if vals else 0
    res[name] = {str(r): [int(x) - base for x in t[r] if x > 0] for r in range(5)}
json.dump(res, open(f"gpurun_out/timeline_r02_{'bf16' if dt == torch.bfloat16 else 'fp32'}.json", "w"))
print({n: {r: len(s) for r, s in d.items()} for n, d in res.items()})
