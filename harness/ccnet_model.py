"""ResNet-101 + RCCA segmentation network with the layer names of networks/ccnet.py, written against stock torch modules.

Why it exists: BASELINE configs[2] / [3] (ResNet101+RCCA forward on 769x769, DDP train step) have to run on the GPU box, where
the reference tree is not mounted, so the network around the operator is restated here.  `tests/test_harness_model.py`
checks in the build container that this definition and the reference's `Seg_Model` (imported unchanged, networks/ccnet.py:
199-205) have identical state-dict keys and shapes, i.e. that released checkpoints address the same tensors.
The backbone is stock torch conv / BatchNorm (north_star); InPlaceABN(Sync) = BatchNorm + leaky ReLU(0.01) (`ABN` below,
convertible to SyncBatchNorm); the head's attention is `cc_attention.CrissCrossAttention` of this repository."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from cc_attention import CrissCrossAttention


class ABN(nn.Module):
    """BatchNorm followed by leaky ReLU -- what InPlaceABN computes (networks/ccnet.py:16, 103-113).  Parameters live directly
    on this module (weight, bias, running_*) like the reference's, so state-dict keys line up."""

    def __init__(self, channels: int, slope: float = 0.01):
        super().__init__()
        self.bn = nn.BatchNorm2d(channels)
        self.slope = slope

    # expose the BatchNorm tensors under this module's own name (keys '<name>.weight', not '<name>.bn.weight')
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        self.bn._save_to_state_dict(destination, prefix, keep_vars)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        if destination is None:
            destination = {}
        self.bn._save_to_state_dict(destination, prefix, keep_vars)
        return destination

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self.bn._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def forward(self, x):
        return F.leaky_relu(self.bn(x), self.slope, inplace=True)


def _conv3x3(cin, cout, stride=1, dilation=1, bias=False):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=dilation, dilation=dilation, bias=bias)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv3x3(planes, planes, stride, dilation)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)), inplace=True)


class RCCAModule(nn.Module):
    """networks/ccnet.py:99-121: conva -> R x criss-cross attention (same module) -> convb -> concat -> classifier."""

    def __init__(self, in_channels, out_channels, num_classes):
        super().__init__()
        inter = in_channels // 4
        self.conva = nn.Sequential(_conv3x3(in_channels, inter), ABN(inter))
        self.cca = CrissCrossAttention(inter)
        self.convb = nn.Sequential(_conv3x3(inter, inter), ABN(inter))
        self.bottleneck = nn.Sequential(_conv3x3(in_channels + inter, out_channels), ABN(out_channels), nn.Dropout2d(0.1),
                                        nn.Conv2d(out_channels, num_classes, 1))

    def forward(self, x, recurrence=1):
        y = self.conva(x)
        for _ in range(recurrence):
            y = self.cca(y)
        y = self.convb(y)
        return self.bottleneck(torch.cat([x, y], 1))


class CCNet(nn.Module):
    """networks/ccnet.py:123-197 (ResNet with deep stem, layers 3/4 dilated 2/4, RCCA head, deep-supervision branch)."""

    def __init__(self, num_classes=19, layers=(3, 4, 23, 3), recurrence=2):
        super().__init__()
        self.inplanes = 128
        self.conv1 = _conv3x3(3, 64, stride=2); self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = _conv3x3(64, 64); self.bn2 = nn.BatchNorm2d(64)
        self.conv3 = _conv3x3(64, 128); self.bn3 = nn.BatchNorm2d(128)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1, ceil_mode=True)
        self.layer1 = self._layer(64, layers[0])
        self.layer2 = self._layer(128, layers[1], stride=2)
        self.layer3 = self._layer(256, layers[2], dilation=2)
        self.layer4 = self._layer(512, layers[3], dilation=4)
        self.head = RCCAModule(2048, 512, num_classes)
        self.dsn = nn.Sequential(_conv3x3(1024, 512, bias=True), ABN(512), nn.Dropout2d(0.1), nn.Conv2d(512, num_classes, 1))
        self.recurrence = recurrence

    def _layer(self, planes, blocks, stride=1, dilation=1):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))
        mods = [Bottleneck(self.inplanes, planes, stride, dilation, down)]
        self.inplanes = planes * 4
        mods += [Bottleneck(self.inplanes, planes, 1, dilation) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = F.relu(self.bn3(self.conv3(x)))
        x = self.layer3(self.layer2(self.layer1(self.maxpool(x))))
        aux = self.dsn(x)
        x = self.head(self.layer4(x), self.recurrence)
        return [x, aux]


def dsn_loss(preds, target, ignore_index=255):
    """loss/criterion.py:10-34 CriterionDSN: cross entropy on both heads at label resolution, 1.0 / 0.4 weights."""
    h, w = target.shape[1:]
    up = lambda p: F.interpolate(p, size=(h, w), mode="bilinear", align_corners=True)
    return (F.cross_entropy(up(preds[0]), target, ignore_index=ignore_index)
            + 0.4 * F.cross_entropy(up(preds[1]), target, ignore_index=ignore_index))
