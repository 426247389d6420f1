"""Pin the oracle (oracle/) against the reference's own outputs (tests/golden/).

The reference has no tests or golden vectors of its own; the fixtures were
produced by running /root/reference/cc_attention/functions.py itself
(tests/golden/make_golden.py).  Everything here runs on CPU.
"""
import numpy as np
import pytest
import torch

from oracle import cca_oracle as O

T = torch.from_numpy


def test_golden_present():
    from conftest import golden_names
    names = golden_names()
    assert len(names) >= 7, names


def test_op_einsum_vs_reference(golden):
    q, k, v = T(golden["q"]), T(golden["k"]), T(golden["v"])
    out, lse = O.cca_forward(q.double(), k.double(), v.double())
    # reference fp64 run is the yardstick; q/k/v were produced in fp32 by the reference
    ref = T(golden["o"]).double()
    assert torch.isfinite(out).all() and torch.isfinite(lse).all()
    assert (out - ref).abs().max().item() < 2e-5, golden["name"]


def test_module_port_vs_reference_fwd_bwd(golden):
    x = T(golden["x"]).clone().requires_grad_(True)
    C = x.shape[1]
    m = O.CrissCrossAttentionOracle(C)
    m.load_state_dict({n[2:]: T(a) for n, a in golden.items() if n.startswith("p_")})
    y = O.rcca_forward(m, x, int(golden["R"]))
    (y * T(golden["g"])).sum().backward()
    assert (y - T(golden["y"])).abs().max().item() < 1e-5
    assert (x.grad - T(golden["dx"])).abs().max().item() < 1e-4
    for n, p in m.named_parameters():
        ref = T(golden["d_" + n])
        tol = 1e-4 * max(1.0, ref.abs().max().item())
        assert (p.grad - ref).abs().max().item() < tol, n


def test_module_port_fp64_vs_reference_fp64(golden):
    x = T(golden["x"]).double()
    m = O.CrissCrossAttentionOracle(x.shape[1]).double()
    m.load_state_dict({n[2:]: T(a).double() for n, a in golden.items() if n.startswith("p_")})
    with torch.no_grad():
        y = O.rcca_forward(m, x, int(golden["R"]))
    assert (y - T(golden["y64"])).abs().max().item() < 1e-12


def test_c_oracle_vs_einsum_and_reference(golden):
    q, k, v = (T(golden[n]).double() for n in "qkv")
    out_c, lse_c = O.cca_forward_c(q, k, v)
    out_e, lse_e = O.cca_forward(q, k, v)
    assert (out_c - out_e).abs().max().item() < 1e-12
    assert (lse_c - lse_e).abs().max().item() < 1e-12
    assert (out_c - T(golden["o"]).double()).abs().max().item() < 2e-5


def test_bruteforce_small():
    torch.manual_seed(0)
    for (B, Cq, C, H, W) in [(2, 3, 5, 4, 6), (1, 2, 3, 1, 5), (1, 2, 3, 5, 1), (1, 1, 1, 1, 1)]:
        q, k, v = torch.randn(B, Cq, H, W).double(), torch.randn(B, Cq, H, W).double(), torch.randn(B, C, H, W).double()
        ob, lb = O.cca_forward_bruteforce(q, k, v)
        oe, le = O.cca_forward(q, k, v)
        assert (ob - oe).abs().max().item() < 1e-12
        assert (lb - le).abs().max().item() < 1e-12


def test_closed_form_backward_vs_autograd_and_c():
    torch.manual_seed(1)
    B, Cq, C, H, W = 2, 4, 6, 5, 7
    q, k, v = (torch.randn(B, c, H, W, dtype=torch.float64, requires_grad=True) for c in (Cq, Cq, C))
    dout = torch.randn(B, C, H, W, dtype=torch.float64)
    out, _ = O.cca_forward(q, k, v)
    gq, gk, gv = torch.autograd.grad(out, (q, k, v), dout)
    dq, dk, dv = O.cca_backward(dout, q.detach(), k.detach(), v.detach())
    cq, ck, cv = O.cca_backward_c(dout, q.detach(), k.detach(), v.detach())
    for a, b_, c in ((gq, dq, cq), (gk, dk, ck), (gv, dv, cv)):
        assert (a - b_).abs().max().item() < 1e-12
        assert (a - c).abs().max().item() < 1e-12


def test_softmax_rows_sum_to_one_and_mask_exact_zero():
    torch.manual_seed(2)
    q, k = torch.randn(1, 4, 6, 5).double(), torch.randn(1, 4, 6, 5).double()
    e_h, e_w = O.cca_logits(q, k)
    a = torch.softmax(torch.cat([e_h, e_w], 3), 3)
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)))
    idx = torch.arange(6)
    assert (a[0, idx, :, idx] == 0).all()          # masked self entry of the column branch


@pytest.mark.parametrize("shape,qt,kt", [((2, 4, 8, 5, 6), 2, 3), ((1, 8, 16, 9, 7), 4, 4), ((1, 2, 4, 1, 5), 1, 2),
                                         ((1, 2, 4, 6, 1), 3, 1), ((1, 4, 8, 13, 11), 16, 5)])
def test_tiled_reduce_add_restatement_equals_the_oracle(shape, qt, kt):
    """The (direction, query tile, key block) decomposition planned for the next kernel generation: final lse
    first, then order-free normalised contributions added onto a zeroed output.  Exact up to fp64 rounding."""
    B, Cq, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + qt + kt)
    q = torch.randn(B, Cq, H, W, generator=g, dtype=torch.float64)
    k = torch.randn(B, Cq, H, W, generator=g, dtype=torch.float64)
    v = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    ro, rl = O.cca_forward(q, k, v)
    to, tl = O.cca_forward_tiled(q, k, v, qt, kt)
    assert (to - ro).abs().max().item() <= 1e-12
    assert (tl - rl).abs().max().item() <= 1e-12
