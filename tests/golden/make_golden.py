"""Generate golden vectors by running the UNMODIFIED reference module on CPU.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports ``/root/reference/cc_attention/functions.py`` as-is.  The reference hard-codes
``.cuda()`` in ``INF`` (functions.py:12); the module stores the function handle on the
instance (functions.py:23), so we override that *instance attribute* with a CPU version
of the same expression -- no reference file is edited or copied.

Each fixture ``cca_<name>.npz`` holds, for one seeded case:
  x, state (7 reference parameters), y = module(x) applied R times, g (upstream grad),
  dx and the 7 parameter grads of ``(y*g).sum()``, plus op-level q, k, v, o of the
  first step (q/k/v = the module's own conv outputs; o = (y1 - x)/gamma).
All in fp32, computed by the reference in fp32 (and y64/o64: the same module in fp64).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location(
        "ref_cc_functions", os.path.join(REF, "cc_attention", "functions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_inf(dtype):
    def INF(B, H, W):  # functions.py:12 without .cuda(), dtype-matched
        return -torch.diag(torch.tensor(float("inf"), dtype=dtype).repeat(H), 0).unsqueeze(0).repeat(B * W, 1, 1)
    return INF


CASES = [
    # name, B, in_dim, H, W, R, gamma, seed, input scale
    ("smoke_2x64x5x6", 2, 64, 5, 6, 1, 1.0, 0, 1.0),        # the reference's own __main__ shape (functions.py:54-55)
    ("c1_1x64x32x32", 1, 64, 32, 32, 1, 1.0, 1, 1.0),       # BASELINE.json configs[0]
    ("r2_2x32x9x7", 2, 32, 9, 7, 2, 0.75, 2, 1.5),          # recurrence 2, H != W
    ("h1_1x16x1x11", 1, 16, 1, 11, 1, 1.0, 3, 1.0),         # H == 1 (column branch fully masked)
    ("w1_1x16x13x1", 1, 16, 13, 1, 1, 1.0, 4, 1.0),         # W == 1
    ("p1_1x8x1x1", 1, 8, 1, 1, 1, 1.0, 5, 1.0),             # single pixel
    ("wide_1x128x17x33", 1, 128, 17, 33, 2, 0.5, 6, 2.0),   # peaky softmax (input scale 2)
]


def run_case(ref, name, B, C, H, W, R, gamma, seed, scale):
    torch.manual_seed(seed)
    m = ref.CrissCrossAttention(C)
    m.INF = cpu_inf(torch.float32)
    with torch.no_grad():
        m.gamma.fill_(gamma)
    x = (torch.randn(B, C, H, W) * scale).requires_grad_(True)
    g = torch.randn(B, C, H, W)
    y = x
    for _ in range(R):                      # networks/ccnet.py:118-119
        y = m(y)
    (y * g).sum().backward()
    out = {"x": x.detach().numpy(), "y": y.detach().numpy(), "g": g.numpy(),
           "dx": x.grad.numpy(), "R": np.int64(R), "gamma": np.float32(gamma)}
    for n, p in m.named_parameters():
        out["p_" + n] = p.detach().numpy()
        out["d_" + n] = p.grad.numpy()
    with torch.no_grad():
        q, k, v = m.query_conv(x), m.key_conv(x), m.value_conv(x)
        y1 = m(x)
        out.update(q=q.numpy(), k=k.numpy(), v=v.numpy(), o=((y1 - x) / gamma).numpy())
        # same module in fp64 (arbiter for fp32 rounding noise)
        m64 = ref.CrissCrossAttention(C).double()
        m64.load_state_dict({n: p.double() for n, p in m.state_dict().items()})
        m64.INF = cpu_inf(torch.float64)
        y64 = x.detach().double()
        for _ in range(R):
            y64 = m64(y64)
        out["y64"] = y64.numpy()
        out["o64"] = ((m64(x.detach().double()) - x.detach().double()) / gamma).numpy()
    np.savez_compressed(os.path.join(HERE, f"cca_{name}.npz"), **out)
    print(name, {k_: v_.shape for k_, v_ in out.items() if hasattr(v_, "shape") and v_.ndim})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    torch.set_num_threads(1)                # deterministic reduction order
    ref = load_reference()
    for case in CASES:
        run_case(ref, *case)
