"""Stand-in for the `inplace_abn` package that networks/ccnet.py:16 imports (a native CUDA extension the reference
depends on but does not vendor).  Numerically InPlaceABN(Sync) is BatchNorm followed by an activation (leaky ReLU 0.01 by
default, 'identity' where ccnet.py:17 asks for it); the in-place memory trick is irrelevant here.  Lives in the harness,
never in the reference tree; `Sync` maps to torch's SyncBatchNorm conversion done by the harness, not here."""
import torch.nn as nn
import torch.nn.functional as F


class InPlaceABN(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, activation="leaky_relu", activation_param=0.01,
                 **kwargs):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self.activation = activation
        self.activation_param = activation_param

    def forward(self, x):
        y = super().forward(x)
        if self.activation == "leaky_relu":
            return F.leaky_relu(y, self.activation_param)
        if self.activation == "relu":
            return F.relu(y)
        if self.activation == "elu":
            return F.elu(y, self.activation_param)
        return y


class InPlaceABNSync(InPlaceABN):
    pass


ABN = InPlaceABN
