"""CPU tests: the C-ABI library loads and exports every symbol of include/cca_b200.h, the
nn.Module mirrors the reference surface, and nothing silently falls back to CPU."""
import ctypes
import os
import re

import pytest
import torch

import cc_attention
import ccnet_b200
from ccnet_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cca_b200.h")).read()
    declared = set(re.findall(r"CCA_API[^;(]*?\b(cca_b200_\w+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_workspace_and_strerror_without_gpu():
    lib = capi.load()
    assert lib.cca_b200_version() == 200
    px = 8 * 97 * 97
    # forward: two partial-lse planes (row + column) + 8 zero-ahead counters; backward: delta + 3 x 8 counters
    assert lib.cca_b200_workspace_bytes(capi.CCA_WS_FORWARD, 8, 64, 512, 97, 97, capi.CCA_F32) == px * 8 + 32
    assert lib.cca_b200_workspace_bytes(capi.CCA_WS_BACKWARD, 8, 64, 512, 97, 97, capi.CCA_F32) == px * 4 + 96
    # lines longer than 112 pixels are tiled: one plane per key block
    assert lib.cca_b200_workspace_bytes(capi.CCA_WS_FORWARD, 1, 64, 512, 193, 193, capi.CCA_F32) == 193 * 193 * 16 + 16
    assert lib.cca_b200_tc_supported(capi.CCA_WS_FORWARD, 1, 8, 64, 32, 32, capi.CCA_F32) == 0      # Cq < 16
    assert lib.cca_b200_strerror(0) == b"ok"
    assert b"unsupported" in lib.cca_b200_strerror(-2)


def test_invalid_arguments_are_rejected_before_any_cuda_call():
    lib = capi.load()
    rc = lib.cca_b200_forward(None, None, None, None, None, None, 0, 1, 8, 64, 4, 4, 0, 0, None)
    assert rc == -1 and b"null" in lib.cca_b200_last_error()
    rc = lib.cca_b200_forward(None, None, None, None, None, None, 0, 0, 8, 64, 4, 4, 0, 0, None)
    assert rc == -1
    rc = lib.cca_b200_forward(None, None, None, None, None, None, 0, 1, 8, 64, 4, 4, 7, 0, None)
    assert rc == -1 and b"dtype" in lib.cca_b200_last_error()


def test_module_surface_matches_reference():
    m = cc_attention.CrissCrossAttention(512)
    assert isinstance(m, ccnet_b200.CrissCrossAttention)
    shapes = {n: tuple(p.shape) for n, p in m.named_parameters()}
    assert shapes == {
        "gamma": (1,),
        "query_conv.weight": (64, 512, 1, 1), "query_conv.bias": (64,),
        "key_conv.weight": (64, 512, 1, 1), "key_conv.bias": (64,),
        "value_conv.weight": (512, 512, 1, 1), "value_conv.bias": (512,),
    }
    assert float(m.gamma) == 0.0                       # functions.py:24
    assert sum(p.numel() for p in m.parameters()) == 328321  # 2*(64*512+64) + 512*512+512 + 1
    assert hasattr(m, "softmax") and hasattr(m, "INF")  # functions.py:22-23


def test_state_dict_roundtrip_with_oracle_module():
    from oracle.cca_oracle import CrissCrossAttentionOracle
    ref = CrissCrossAttentionOracle(64)
    m = cc_attention.CrissCrossAttention(64)
    missing, unexpected = m.load_state_dict(ref.state_dict(), strict=False)
    assert not missing and not unexpected


def test_no_cpu_fallback():
    m = cc_attention.CrissCrossAttention(64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 64, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        ccnet_b200.cca_forward(torch.randn(1, 8, 4, 4), torch.randn(1, 8, 4, 4), torch.randn(1, 64, 4, 4))


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "ccnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    assert "oracle" not in open(os.path.join(ROOT, "cc_attention", "__init__.py")).read().replace("oracle/", "")


def test_torch_ops_are_registered_with_fake_implementations():
    """torch.ops.cca.{forward, backward, forward_residual} exist and shape-infer under FakeTensorMode, so a
    traced / compiled networks/ccnet.py does not graph-break on the operator."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    assert all(hasattr(torch.ops.cca, n) for n in ("forward", "backward", "forward_residual"))
    with FakeTensorMode():
        q = torch.empty(2, 8, 5, 6, device="cuda")
        v = torch.empty(2, 64, 5, 6, device="cuda")
        out, lse = torch.ops.cca.forward(q, q, v)
        assert out.shape == v.shape and out.dtype == v.dtype and lse.shape == (2, 5, 6) and lse.dtype == torch.float32
        dq, dk, dv = torch.ops.cca.backward(out, q, q, v, out, lse)
        assert dq.shape == q.shape and dv.shape == v.shape
        y, lse2, o2 = torch.ops.cca.forward_residual(q, q, v, v, torch.empty(1, device="cuda"))
        assert y.shape == v.shape and lse2.shape == lse.shape and o2.shape == v.shape
    torch.library.opcheck  # noqa: B018  (present in this torch; the GPU suite runs it on real tensors)
