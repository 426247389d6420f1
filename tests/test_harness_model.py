"""The harness network (harness/ccnet_model.py, used for BASELINE configs[2] / [3] on the GPU box) names and shapes every tensor
exactly like the reference's Seg_Model (networks/ccnet.py imported unchanged with the inplace_abn stand-in)."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "networks", "ccnet.py")), reason="reference mount absent")
def test_harness_network_matches_reference_state_dict():
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path[:0] = [ROOT, os.path.join(ROOT, "harness", "shims"), REF]
    try:
        ref_ccnet = importlib.import_module("networks.ccnet")
        with torch.device("meta"):
            ref = ref_ccnet.Seg_Model(num_classes=19, recurrence=2)
            from harness.ccnet_model import CCNet
            ours = CCNet(num_classes=19, recurrence=2)
        a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert a == b, (sorted(set(a) ^ set(b))[:20], [k for k in a if k in b and a[k] != b[k]][:10])
        assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters())
    finally:
        sys.path[:] = saved_path
        for name in list(sys.modules):
            if name not in saved_mods and (name.startswith("networks") or name.startswith("utils") or name == "inplace_abn"):
                del sys.modules[name]


def test_harness_network_forward_shapes_on_cpu_fallback_free():
    """Tiny structural check without the operator: the head's attention needs CUDA, so only the backbone + dsn run here."""
    from harness.ccnet_model import CCNet
    net = CCNet(num_classes=5, layers=(1, 1, 1, 1), recurrence=1).eval()
    x = torch.randn(1, 3, 65, 65)
    with torch.no_grad():
        y = torch.relu(net.bn1(net.conv1(x)))
        y = torch.relu(net.bn2(net.conv2(y)))
        y = net.maxpool(torch.relu(net.bn3(net.conv3(y))))
        y = net.layer3(net.layer2(net.layer1(y)))
        assert y.shape == (1, 1024, 9, 9)                # 65 -> 33 -> 17 -> 9: output stride 8
        assert net.dsn(y).shape == (1, 5, 9, 9)
        assert net.layer4(y).shape == (1, 2048, 9, 9)
