"""CPU oracle for criss-cross attention (TEST INFRASTRUCTURE - not a product path).

This file is a CPU restatement of the reference hot path
``/root/reference/cc_attention/functions.py:27-49`` (``CrissCrossAttention.forward``)
and of its autograd backward.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``ccnet_b200`` / ``cc_attention``) never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests / golden vectors, so the
oracle is pinned against outputs of the reference module itself, generated in the build
container by ``tests/golden/make_golden.py`` (imports ``/root/reference/cc_attention``
unmodified, overriding only the instance attribute ``INF`` so it runs on CPU) and
committed under ``tests/golden/``.  ``tests/test_oracle.py`` checks every function here
against those fixtures.

Three independent restatements live here:
  * ``cca_forward`` / ``cca_backward``  - einsum, any float dtype (fp64 = arbiter)
  * ``cca_forward_bruteforce``          - pure-python triple loop (tiny maps only)
  * ``CrissCrossAttentionOracle``       - module-level port (1x1 convs + op + residual),
                                          same parameter names as the reference.
A fourth one, in plain C, is ``oracle/cca_oracle.c`` (``load_c_oracle()`` binds it).
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# op level:  (q[B,Cq,H,W], k[B,Cq,H,W], v[B,C,H,W]) -> out[B,C,H,W], lse[B,H,W]
# --------------------------------------------------------------------------------------
def cca_logits(q: torch.Tensor, k: torch.Tensor):
    """Column- and row-branch affinities.

    e_h[b,h,w,g] = sum_c q[b,c,h,w] * k[b,c,g,w]  with -inf at g == h
        (functions.py:38: bmm(proj_query_H, proj_key_H) + INF, then view/permute)
    e_w[b,h,w,g] = sum_c q[b,c,h,w] * k[b,c,h,g]
        (functions.py:39)
    """
    B, _, H, W = q.shape
    e_h = torch.einsum("bchw,bcgw->bhwg", q, k)
    eye = torch.eye(H, dtype=torch.bool, device=q.device).view(1, H, 1, H)
    e_h = e_h.masked_fill(eye, float("-inf"))
    e_w = torch.einsum("bchw,bchg->bhwg", q, k)
    return e_h, e_w


def cca_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
    """One criss-cross step at op level (functions.py:38-47), returns (out, lse).

    out = out_H + out_W  (the ``gamma*(...)+x`` residual of functions.py:49 is the
    caller's job); lse[b,h,w] = logsumexp over the H+W logits (one of them -inf).
    """
    H = q.shape[2]
    e_h, e_w = cca_logits(q, k)
    e = torch.cat([e_h, e_w], dim=3)                      # functions.py:40 (cat)
    lse = torch.logsumexp(e, dim=3)
    a = torch.softmax(e, dim=3)                           # functions.py:40 (Softmax(dim=3))
    a_h, a_w = a[..., :H], a[..., H:]                     # functions.py:42,45
    out_h = torch.einsum("bhwg,bcgw->bchw", a_h, v)       # functions.py:46
    out_w = torch.einsum("bhwg,bchg->bchw", a_w, v)       # functions.py:47
    return out_h + out_w, lse


def cca_backward(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
    """Closed-form gradient of ``cca_forward`` w.r.t. q, k, v.

    dA = dO.V, delta = sum_j A_j dA_j, dE = A*(dA-delta), then the four bmm transposes.
    """
    H = q.shape[2]
    e_h, e_w = cca_logits(q, k)
    a = torch.softmax(torch.cat([e_h, e_w], dim=3), dim=3)
    a_h, a_w = a[..., :H], a[..., H:]
    da_h = torch.einsum("bchw,bcgw->bhwg", dout, v)
    da_w = torch.einsum("bchw,bchg->bhwg", dout, v)
    delta = (a_h * da_h).sum(-1, keepdim=True) + (a_w * da_w).sum(-1, keepdim=True)
    de_h = a_h * (da_h - delta)
    de_w = a_w * (da_w - delta)
    dv = torch.einsum("bhwg,bchw->bcgw", a_h, dout) + torch.einsum("bhwg,bchw->bchg", a_w, dout)
    dq = torch.einsum("bhwg,bcgw->bchw", de_h, k) + torch.einsum("bhwg,bchg->bchw", de_w, k)
    dk = torch.einsum("bhwg,bchw->bcgw", de_h, q) + torch.einsum("bhwg,bchw->bchg", de_w, q)
    return dq, dk, dv


def cca_forward_tiled(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, qt: int, kt: int):
    """The same criss-cross step evaluated the way DESIGN.md 8 plans the next kernel generation: independent
    (direction, query tile, key block) items.  Phase 1 folds per-item row maxima / sums into the final lse (only q, k);
    phase 2 lets every item ADD exp(S_item - lse) @ V_item onto a zeroed output -- no partial-output merge chain, any item
    order, tiles of at most qt queries x kt keys along a line.  Must equal cca_forward (tests/test_oracle.py); it is the
    restatement the tiled CUDA path will be checked against."""
    B, _, H, W = q.shape
    C = v.shape[1]
    neg = float("-inf")
    items = []                                            # (direction, line index, query range, key range)
    for w in range(W):
        for q0 in range(0, H, qt):
            for k0 in range(0, H, kt):
                items.append(("col", w, (q0, min(q0 + qt, H)), (k0, min(k0 + kt, H))))
    for h in range(H):
        for q0 in range(0, W, qt):
            for k0 in range(0, W, kt):
                items.append(("row", h, (q0, min(q0 + qt, W)), (k0, min(k0 + kt, W))))

    def logits(it):
        d, i, (a, b), (c, e) = it
        if d == "col":
            s = torch.einsum("bcq,bcg->bqg", q[:, :, a:b, i], k[:, :, c:e, i])
            qi = torch.arange(a, b).view(-1, 1)
            ki = torch.arange(c, e).view(1, -1)
            return s.masked_fill((qi == ki).unsqueeze(0), neg)             # the self position of the column branch
        return torch.einsum("bcq,bcg->bqg", q[:, :, i, a:b], k[:, :, i, c:e])

    # phase 1: running (m, l) per pixel -> lse
    m = torch.full((B, H, W), neg, dtype=q.dtype)
    l = torch.zeros((B, H, W), dtype=q.dtype)
    for it in items:
        d, i, (a, b), _ = it
        s = logits(it)
        mi = s.max(dim=2).values
        sl = (slice(None), slice(a, b), i) if d == "col" else (slice(None), i, slice(a, b))
        mo, lo = m[sl], l[sl]
        mn = torch.maximum(mo, mi)
        safe = torch.where(torch.isinf(mn), torch.zeros_like(mn), mn)      # an all-masked item (1x1 column tile)
        l[sl] = lo * torch.exp(mo - safe) + torch.exp(s - safe.unsqueeze(2)).sum(dim=2)
        m[sl] = mn
    lse = m + torch.log(l)
    # phase 2: every item adds its normalised contribution (order-free)
    out = torch.zeros((B, C, H, W), dtype=v.dtype)
    for it in reversed(items):                                               # any order
        d, i, (a, b), (c, e) = it
        s = logits(it)
        if d == "col":
            p = torch.exp(s - lse[:, a:b, i].unsqueeze(2))
            out[:, :, a:b, i] += torch.einsum("bqg,bcg->bcq", p, v[:, :, c:e, i])
        else:
            p = torch.exp(s - lse[:, i, a:b].unsqueeze(2))
            out[:, :, i, a:b] += torch.einsum("bqg,bcg->bcq", p, v[:, :, i, c:e])
    return out, lse


def cca_forward_bruteforce(q, k, v):
    """Pure-python loops over the criss-cross set of every pixel (tiny inputs only).

    Independent of einsum index conventions: for pixel (h,w) the set is
    {(g,w): g != h} U {(h,g): all g}  -> H+W-1 entries (functions.py:38-40).
    """
    B, Cq, H, W = q.shape
    C = v.shape[1]
    qd, kd, vd = q.double(), k.double(), v.double()
    out = torch.zeros(B, C, H, W, dtype=torch.float64)
    lse = torch.zeros(B, H, W, dtype=torch.float64)
    for b in range(B):
        for h in range(H):
            for w in range(W):
                pos = [(g, w) for g in range(H) if g != h] + [(h, g) for g in range(W)]
                logits = [float((qd[b, :, h, w] * kd[b, :, y, x]).sum()) for (y, x) in pos]
                m = max(logits)
                ex = [math.exp(t - m) for t in logits]
                s = sum(ex)
                lse[b, h, w] = m + math.log(s)
                for (y, x), e in zip(pos, ex):
                    out[b, :, h, w] += (e / s) * vd[b, :, y, x]
    return out, lse


# --------------------------------------------------------------------------------------
# module level port (used as the CPU baseline "port" and for module parity)
# --------------------------------------------------------------------------------------
class CrissCrossAttentionOracle(nn.Module):
    """Module-level port of functions.py:15-49 with the same parameter names.

    Uses the same torch ops as the reference (Conv2d 1x1, bmm, cat, softmax) so that its
    CPU timing is a fair stand-in for the reference module (which cannot travel to the
    GPU box); the only change is that the INF mask is built on the input's device/dtype
    (the reference hard-codes .cuda(), functions.py:12).
    """

    def __init__(self, in_dim: int):
        super().__init__()
        self.query_conv = nn.Conv2d(in_dim, in_dim // 8, kernel_size=1)   # functions.py:19
        self.key_conv = nn.Conv2d(in_dim, in_dim // 8, kernel_size=1)     # functions.py:20
        self.value_conv = nn.Conv2d(in_dim, in_dim, kernel_size=1)        # functions.py:21
        self.gamma = nn.Parameter(torch.zeros(1))                         # functions.py:24

    @staticmethod
    def _inf(B, H, W, like):
        d = torch.diag(torch.full((H,), float("inf"), dtype=like.dtype, device=like.device))
        return -d.unsqueeze(0).repeat(B * W, 1, 1)                        # functions.py:12

    def forward(self, x):
        B, _, H, W = x.size()
        q = self.query_conv(x)
        q_h = q.permute(0, 3, 1, 2).contiguous().view(B * W, -1, H).permute(0, 2, 1)
        q_w = q.permute(0, 2, 1, 3).contiguous().view(B * H, -1, W).permute(0, 2, 1)
        k = self.key_conv(x)
        k_h = k.permute(0, 3, 1, 2).contiguous().view(B * W, -1, H)
        k_w = k.permute(0, 2, 1, 3).contiguous().view(B * H, -1, W)
        v = self.value_conv(x)
        v_h = v.permute(0, 3, 1, 2).contiguous().view(B * W, -1, H)
        v_w = v.permute(0, 2, 1, 3).contiguous().view(B * H, -1, W)
        e_h = (torch.bmm(q_h, k_h) + self._inf(B, H, W, q)).view(B, W, H, H).permute(0, 2, 1, 3)
        e_w = torch.bmm(q_w, k_w).view(B, H, W, W)
        a = torch.softmax(torch.cat([e_h, e_w], 3), dim=3)
        a_h = a[:, :, :, 0:H].permute(0, 2, 1, 3).contiguous().view(B * W, H, H)
        a_w = a[:, :, :, H:H + W].contiguous().view(B * H, W, W)
        o_h = torch.bmm(v_h, a_h.permute(0, 2, 1)).view(B, W, -1, H).permute(0, 2, 3, 1)
        o_w = torch.bmm(v_w, a_w.permute(0, 2, 1)).view(B, H, -1, W).permute(0, 2, 1, 3)
        return self.gamma * (o_h + o_w) + x                               # functions.py:49


def rcca_forward(module: nn.Module, x: torch.Tensor, recurrence: int) -> torch.Tensor:
    """The recurrence loop of networks/ccnet.py:118-119 (same module R times)."""
    out = x
    for _ in range(recurrence):
        out = module(out)
    return out


# --------------------------------------------------------------------------------------
# binding of the plain-C restatement (oracle/cca_oracle.c -> oracle/libcca_oracle.so)
# --------------------------------------------------------------------------------------
def c_oracle_path() -> str:
    return os.path.join(_HERE, "libcca_oracle.so")


def build_c_oracle(force: bool = False) -> str:
    import subprocess

    so = c_oracle_path()
    src = os.path.join(_HERE, "cca_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", so, src, "-lm"], cwd=_HERE
        )
    return so


def load_c_oracle():
    lib = ctypes.CDLL(build_c_oracle())
    dp = ctypes.POINTER(ctypes.c_double)
    lib.cca_oracle_forward_f64.argtypes = [dp, dp, dp, dp, dp] + [ctypes.c_int] * 5
    lib.cca_oracle_forward_f64.restype = None
    lib.cca_oracle_backward_f64.argtypes = [dp] * 7 + [ctypes.c_int] * 5
    lib.cca_oracle_backward_f64.restype = None
    return lib


def _dptr(t: torch.Tensor):
    assert t.dtype == torch.float64 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_double))


def cca_forward_c(q, k, v):
    lib = load_c_oracle()
    q, k, v = (t.double().contiguous() for t in (q, k, v))
    B, Cq, H, W = q.shape
    C = v.shape[1]
    out = torch.empty(B, C, H, W, dtype=torch.float64)
    lse = torch.empty(B, H, W, dtype=torch.float64)
    lib.cca_oracle_forward_f64(_dptr(q), _dptr(k), _dptr(v), _dptr(out), _dptr(lse), B, Cq, C, H, W)
    return out, lse


def cca_backward_c(dout, q, k, v):
    lib = load_c_oracle()
    dout, q, k, v = (t.double().contiguous() for t in (dout, q, k, v))
    B, Cq, H, W = q.shape
    C = v.shape[1]
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    lib.cca_oracle_backward_f64(_dptr(dout), _dptr(q), _dptr(k), _dptr(v),
                                _dptr(dq), _dptr(dk), _dptr(dv), B, Cq, C, H, W)
    return dq, dk, dv
