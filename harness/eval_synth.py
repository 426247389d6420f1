"""Checkpoint loading and the inference (evaluation) path around the operator -- SURVEY.md 8(f) N4.

What the reference does (utils/pyt_utils.py:47-85 `load_model`; evaluate.py:102-173 `pad_image`, `predict_sliding`,
`predict_whole`, `predict_multiscale`; evaluate.py:175-194 `get_confusion_matrix`; evaluate.py:268-274 mean IoU), restated for
current torch and kept on the device: the reference round-trips every tile through numpy on the host (`.cpu().numpy()`, float64
accumulators, `scipy.ndimage.zoom`); here the probability map, the hit counts and the confusion matrix are device tensors and
only the final numbers leave the GPU.  Same tiling arithmetic, same interpolation (bilinear, align_corners=True), same
averaging; `tests/test_eval_path.py` pins it to fixtures produced by the reference's own functions
(tests/golden/make_eval_golden.py).

  python harness/eval_synth.py [--size 1024,2048 --tile 769,769 --recurrence 2]     one synthetic Cityscapes-sized image
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from collections import OrderedDict
from math import ceil

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ---------------------------------------------------------------------------------------------------------------------
# checkpoints (utils/pyt_utils.py:47-85)
# ---------------------------------------------------------------------------------------------------------------------
def load_model(model, model_file, is_restore: bool = False, map_location="cpu"):
    """`utils/pyt_utils.load_model`: `model_file` is a path or a state dict; a dict with a 'model' entry is unwrapped
    (pyt_utils.py:52-53); `is_restore` prefixes every key with 'module.' (:58-63); loading is non-strict (:65).
    Returns (model, missing_keys, unexpected_keys) -- the reference only logs the two sets (:66-77).

    One addition for checkpoints written from a DataParallel / DDP wrapper (the released CS_scenes_60000.pth was saved from
    `model.module`-less training code, snapshots of this repo's harness are saved from the bare network): when the model has no
    'module.' keys and EVERY checkpoint key has the prefix, it is stripped."""
    if isinstance(model_file, (str, os.PathLike)):
        state = torch.load(model_file, map_location=map_location, weights_only=True)
        if "model" in state:
            state = state["model"]
    else:
        state = model_file
    if is_restore:
        state = OrderedDict(("module." + k, v) for k, v in state.items())
    own = model.state_dict()
    if state and all(k.startswith("module.") for k in state) and not any(k.startswith("module.") for k in own):
        state = OrderedDict((k[len("module."):], v) for k, v in state.items())
    model.load_state_dict(state, strict=False)
    ckpt_keys, own_keys = set(state.keys()), set(own.keys())
    return model, sorted(own_keys - ckpt_keys), sorted(ckpt_keys - own_keys)


# ---------------------------------------------------------------------------------------------------------------------
# inference (evaluate.py:95-173)
# ---------------------------------------------------------------------------------------------------------------------
def _first(pred):
    return pred[0] if isinstance(pred, (list, tuple)) else pred       # evaluate.py:129-130 (the network returns [seg, dsn])


def tile_grid(image_hw, tile_hw, overlap: float = 1 / 3):
    """The windows `predict_sliding` visits (evaluate.py:106-123): stride = ceil(tile_h * (1 - overlap)) in both directions, the
    last window of a row / column is pulled back inside the image.  Returns [(y1, y2, x1, x2)]."""
    H, W = image_hw
    th, tw = tile_hw
    stride = ceil(th * (1 - overlap))
    rows = int(ceil((H - th) / stride) + 1)
    cols = int(ceil((W - tw) / stride) + 1)
    out = []
    for r in range(rows):
        for c in range(cols):
            x1, y1 = int(c * stride), int(r * stride)
            x2, y2 = min(x1 + tw, W), min(y1 + th, H)
            x1, y1 = max(int(x2 - tw), 0), max(int(y2 - th), 0)
            out.append((y1, y2, x1, x2))
    return out


@torch.no_grad()
def predict_sliding(net, image: torch.Tensor, tile_size, classes: int, tile_batch: int = 1):
    """evaluate.py:102-142.  image [N,3,H,W] on the network's device; returns [N,H,W,classes] (float32, same device): every
    window is zero-padded up to the tile (evaluate.py:95-100), run, upsampled to the tile size (bilinear, align_corners=True),
    cropped back and accumulated; overlaps are averaged.  `tile_batch` > 1 stacks windows into one forward (the reference
    runs them one by one; the result is the same up to the batch-invariance of the network's kernels).  For N > 1 every sample
    gets its own prediction (evaluate.py:132 adds sample 0's window to all N; the reference's loader runs N = 1)."""
    N, _, H, W = image.shape
    th, tw = tile_size
    probs = torch.zeros(N, H, W, classes, device=image.device, dtype=torch.float32)
    count = torch.zeros(1, H, W, 1, device=image.device, dtype=torch.float32)
    wins = tile_grid((H, W), (th, tw))
    for i in range(0, len(wins), max(1, tile_batch)):
        group = wins[i:i + max(1, tile_batch)]
        tiles = []
        for (y1, y2, x1, x2) in group:
            t = image[:, :, y1:y2, x1:x2]
            tiles.append(F.pad(t, (0, tw - (x2 - x1), 0, th - (y2 - y1))))
        pred = _first(net(torch.cat(tiles, 0)))
        pred = F.interpolate(pred.float(), size=(th, tw), mode="bilinear", align_corners=True)
        for j, (y1, y2, x1, x2) in enumerate(group):
            p = pred[j * N:(j + 1) * N, :, :y2 - y1, :x2 - x1].permute(0, 2, 3, 1)
            probs[:, y1:y2, x1:x2] += p
            count[:, y1:y2, x1:x2] += 1
    return probs / count


@torch.no_grad()
def predict_whole(net, image: torch.Tensor):
    """evaluate.py:144-152: one forward on the whole image, upsampled back to its size."""
    H, W = image.shape[2:]
    pred = F.interpolate(_first(net(image)).float(), size=(H, W), mode="bilinear", align_corners=True)
    return pred.permute(0, 2, 3, 1)


def zoom_bilinear(image: torch.Tensor, scale: float):
    """`scipy.ndimage.zoom(image, (1, 1, s, s), order=1, prefilter=False)` (evaluate.py:165): output size round(n * s), samples at
    i * (n_in - 1) / (n_out - 1) -- which is bilinear interpolation with align_corners=True."""
    H, W = image.shape[2:]
    oh, ow = int(round(H * scale)), int(round(W * scale))
    if (oh, ow) == (H, W):
        return image
    return F.interpolate(image, size=(oh, ow), mode="bilinear", align_corners=True)


@torch.no_grad()
def predict_multiscale(net, image: torch.Tensor, tile_size, scales, classes: int, flip_evaluation: bool, whole: bool = False,
                       tile_batch: int = 1, unflip_axis: int = 2):
    """evaluate.py:154-173: mean over scales of the sliding-window prediction (optionally averaged with the prediction of the
    mirrored image, mirrored back).  As in the reference each scale's map has that scale's size, so -- as there -- more than one
    scale only adds up when every scale rounds to the same size (the reference evaluates with scales=[1.0], evaluate.py:248).
    `unflip_axis`: the mirrored prediction [N,H,W,classes] is mirrored back along W (axis 2).  evaluate.py:171 writes
    `flip_scaled_probs[:,::-1,:]`, which reverses axis 1 (the rows); pass unflip_axis=1 to reproduce that line bit for bit (the
    reference never takes this branch: its main() calls with flip_evaluation=False)."""
    N, _, H, W = image.shape
    total = None
    for s in scales:
        img = zoom_bilinear(image, float(s))
        run = (lambda x: predict_whole(net, x)) if whole else (lambda x: predict_sliding(net, x, tile_size, classes, tile_batch))
        p = run(img)
        if flip_evaluation:
            p = 0.5 * (p + run(img.flip(3)).flip(unflip_axis))
        total = p if total is None else total + p
    return total / len(scales)


def get_confusion_matrix(gt_label: torch.Tensor, pred_label: torch.Tensor, class_num: int):
    """evaluate.py:175-194: counts[gt, pred] over the given (already ignore-filtered) labels."""
    idx = gt_label.reshape(-1).long() * class_num + pred_label.reshape(-1).long()
    return torch.bincount(idx, minlength=class_num * class_num)[:class_num * class_num].reshape(class_num, class_num).double()


def mean_iou(confusion: torch.Tensor):
    """evaluate.py:268-274."""
    pos, res, tp = confusion.sum(1), confusion.sum(0), confusion.diag()
    iu = tp / torch.clamp(pos + res - tp, min=1.0)
    return iu.mean().item(), iu


@torch.no_grad()
def evaluate_batch(net, image, label, tile_size, classes, ignore_label=255, scales=(1.0,), flip=False, whole=False, tile_batch=1):
    """One iteration of the loop at evaluate.py:246-262 (without the PNG dump): prediction -> argmax -> confusion matrix."""
    probs = predict_multiscale(net, image, tile_size, scales, classes, flip, whole, tile_batch)
    pred = probs.argmax(dim=3)
    keep = label != ignore_label
    return get_confusion_matrix(label[keep], pred[keep], classes)


def run(local_rank: int = 0, size=(1024, 2048), tile=(769, 769), recurrence: int = 2, classes: int = 19, steps: int = 1,
        warmup: int = 1, tile_batch: int = 1, checkpoint: str | None = None):
    """Synthetic analogue of evaluate.py's main(): a Cityscapes-sized random image through ResNet101+RCCA with sliding windows.
    Returns a dict (ms per image, windows, mean IoU against random labels -- a smoke number, not an accuracy)."""
    from harness.ccnet_model import CCNet
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(11)
    net = CCNet(num_classes=classes, recurrence=recurrence).to(dev).to(memory_format=torch.channels_last).eval()
    missing = unexpected = None
    if checkpoint:
        net, missing, unexpected = load_model(net, checkpoint)
    with torch.no_grad():
        net.head.cca.gamma.fill_(0.5)
    image = torch.randn(1, 3, *size, device=dev)
    label = torch.randint(0, classes, (1, *size), device=dev)
    label[:, :16] = 255
    for _ in range(warmup):
        conf = evaluate_batch(net, image, label, tile, classes, tile_batch=tile_batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        conf = evaluate_batch(net, image, label, tile, classes, tile_batch=tile_batch)
    e1.record()
    torch.cuda.synchronize()
    miou, _ = mean_iou(conf)
    out = {"image": list(size), "tile": list(tile), "windows": len(tile_grid(size, tile)), "recurrence": recurrence,
           "tile_batch": tile_batch, "ms_per_image": e0.elapsed_time(e1) / steps, "mean_iou_vs_random_labels": miou}
    if checkpoint:
        out["missing_keys"], out["unexpected_keys"] = len(missing), len(unexpected)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1024,2048")
    ap.add_argument("--tile", default="769,769")
    ap.add_argument("--recurrence", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--tile-batch", type=int, default=1)
    ap.add_argument("--restore-from", default=None)
    a = ap.parse_args()
    print(json.dumps(run(0, tuple(map(int, a.size.split(","))), tuple(map(int, a.tile.split(","))), a.recurrence, 19, a.steps,
                         a.warmup, a.tile_batch, a.restore_from)))
