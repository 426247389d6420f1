"""Fixture for the evaluation path: runs the REFERENCE's own `pad_image`, `predict_sliding`,
`predict_multiscale` and `get_confusion_matrix` (evaluate.py:95-194) on a synthetic image with a small fixed network and stores
inputs and outputs.  evaluate.py cannot be imported as a module here (cv2, the dataset package, an Engine), so the five
function definitions are taken out of its source with `ast` at generation time and executed against numpy / torch / scipy;
`.cuda()` is a no-op in this GPU-less container.  Nothing of the reference is written into the repository -- only the arrays.

    python tests/golden/make_eval_golden.py        (needs /root/reference; writes tests/golden/eval_sliding.npz)
"""
import ast
import os
from math import ceil

import numpy as np
import torch
import torch.nn as nn
from scipy import ndimage

REF = "/root/reference/evaluate.py"
HERE = os.path.dirname(os.path.abspath(__file__))
WANT = ("pad_image", "predict_sliding", "predict_whole", "predict_multiscale", "get_confusion_matrix")


def reference_functions():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    ns = {"np": np, "torch": torch, "nn": nn, "ceil": ceil, "ndimage": ndimage}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def tiny_net(classes):
    torch.manual_seed(5)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, stride=2, padding=1), nn.ReLU(), nn.Conv2d(8, classes, 3, stride=2, padding=1)).double().eval()
    return net


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    fn = reference_functions()
    classes, tile = 5, (32, 32)
    net = tiny_net(classes)

    class ListNet(nn.Module):                       # the reference network returns [seg, dsn] (networks/ccnet.py:141)
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            y = self.m(x)
            return [y, y * 0]

    rng = np.random.default_rng(3)
    image = rng.standard_normal((1, 3, 50, 70))
    out = {"image": image, "classes": classes, "tile": np.array(tile)}
    for k, v in net.state_dict().items():
        out["net." + k] = v.numpy()
    with torch.no_grad():
        out["sliding"] = fn["predict_sliding"](ListNet(net), image, tile, classes, 0)
        out["whole"] = fn["predict_whole"](ListNet(net), image, tile, 0)
        out["multi_flip"] = fn["predict_multiscale"](ListNet(net), torch.from_numpy(image), tile, [1.0], classes, True, 0)
        # (a scale other than 1.0 does not run in the reference: evaluate.py:173 adds maps of different sizes)
        small = rng.standard_normal((1, 3, 20, 45))            # smaller than the tile in one direction: padding path
        out["image_small"] = small
        out["sliding_small"] = fn["predict_sliding"](ListNet(net), small, tile, classes, 0)
    out["zoom075"] = ndimage.zoom(image, (1.0, 1.0, 0.75, 0.75), order=1, prefilter=False)
    gt = rng.integers(0, classes, size=4000)
    pr = rng.integers(0, classes, size=4000)
    out["cm_gt"], out["cm_pred"] = gt, pr
    out["cm"] = fn["get_confusion_matrix"](gt, pr, classes)
    for k in ("sliding", "whole", "multi_flip", "sliding_small", "zoom075"):
        out[k] = out[k].astype(np.float32)              # compared at 1e-5; halves the fixture
    np.savez_compressed(os.path.join(HERE, "eval_sliding.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
