"""``torch.ops.cca.*`` -- the operator ABI SURVEY.md 8(b) specifies, registered with torch's dispatcher:

    torch.ops.cca.forward(q, k, v)                        -> (out, lse)
    torch.ops.cca.backward(dout, q, k, v, out, lse)       -> (dq, dk, dv)
    torch.ops.cca.forward_residual(q, k, v, x, gamma)     -> (y, lse)        y = gamma * out + x  (functions.py:49)

CUDA implementations call the C ABI (ccnet_b200.functional -> libcca_b200.so); FakeTensor ("meta") implementations give
shapes / dtypes / memory formats so that ``torch.compile`` and ``torch.export`` trace through ``networks/ccnet.py`` without a
graph break; autograd is registered on ``forward`` and ``forward_residual``.  Registration happens through ``torch.library``
(the Python face of TORCH_LIBRARY): the kernels themselves stay behind the torch-free C ABI."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import functional as F_


def _out_like(q: Tensor, v: Tensor) -> Tuple[Tensor, Tensor]:
    B, Cq, H, W = q.shape
    fmt = torch.channels_last if F_.tc_eligible(B, Cq, v.shape[1], H, W, q.dtype) else torch.contiguous_format
    return torch.empty(v.shape, dtype=v.dtype, device=v.device).contiguous(memory_format=fmt), \
        torch.empty((B, H, W), dtype=torch.float32, device=q.device)


@torch.library.custom_op("cca::forward", mutates_args=(), device_types="cuda")
def forward(q: Tensor, k: Tensor, v: Tensor) -> Tuple[Tensor, Tensor]:
    return F_.cca_forward(q, k, v)


@forward.register_fake
def _(q, k, v):
    return _out_like(q, v)


@torch.library.custom_op("cca::backward", mutates_args=(), device_types="cuda")
def backward(dout: Tensor, q: Tensor, k: Tensor, v: Tensor, out: Tensor, lse: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    return F_.cca_backward(dout, q, k, v, out, lse)


@backward.register_fake
def _(dout, q, k, v, out, lse):
    fmt = torch.channels_last if out.is_contiguous(memory_format=torch.channels_last) and out.dim() == 4 else torch.contiguous_format
    mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=t.device).contiguous(memory_format=fmt)
    return mk(q), mk(k), mk(v)


def _fwd_setup(ctx, inputs, output):
    q, k, v = inputs
    out, lse = output
    ctx.save_for_backward(q, k, v, out, lse)


def _fwd_backward(ctx, dout, dlse):
    q, k, v, out, lse = ctx.saved_tensors
    dq, dk, dv = torch.ops.cca.backward(dout.contiguous(), q, k, v, out, lse)
    return dq, dk, dv


forward.register_autograd(_fwd_backward, setup_context=_fwd_setup)


@torch.library.custom_op("cca::forward_residual", mutates_args=(), device_types="cuda")
def forward_residual(q: Tensor, k: Tensor, v: Tensor, x: Tensor, gamma: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(y, lse, out): y = gamma * out + x; `out` is returned for the backward (delta = <dy, out>)."""
    out, lse = F_.cca_forward(q, k, v)
    return torch.addcmul(x, gamma, out), lse, out


@forward_residual.register_fake
def _(q, k, v, x, gamma):
    out, lse = _out_like(q, v)
    return torch.empty_like(out), lse, out


def _res_setup(ctx, inputs, output):
    q, k, v, x, gamma = inputs
    y, lse, out = output
    ctx.save_for_backward(q, k, v, out, lse, gamma)


def _res_backward(ctx, dy, dlse, dout_unused):
    q, k, v, out, lse, gamma = ctx.saved_tensors
    # the attention backward is linear in dout: run it on dy and scale the (much smaller / equally sized) results by gamma
    dq, dk, dv = torch.ops.cca.backward(dy.contiguous(), q, k, v, out, lse)
    g = gamma.to(dy.dtype)
    return dq * g, dk * g, dv * g, dy, (dy * out).sum().reshape(gamma.shape).to(gamma.dtype)


forward_residual.register_autograd(_res_backward, setup_context=_res_setup)
