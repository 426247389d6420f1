"""CPU tests of the tensor-core kernels' work decomposition (ccnet_b200/csrc/cca_items.cuh) through the host-only
introspection entry points of the C ABI: every (query pixel, key pixel) pair of every line is owned by exactly one item,
tiles fit the kernel templates, the order is dependency-safe, the zero-ahead shares tile a sample exactly."""
import ctypes

import numpy as np
import pytest

from ccnet_b200 import capi

SHAPES = [(1, 1, 1), (2, 5, 6), (1, 97, 97), (8, 97, 97), (2, 65, 65), (1, 112, 80), (1, 113, 200), (1, 129, 129),
          (1, 193, 193), (2, 20, 97), (1, 1, 300), (1, 896, 17)]


def _space(B, H, W):
    out = (ctypes.c_int * 8)()
    capi.load().cca_b200_item_space(B, H, W, out)
    return dict(zip(("total", "per_sample", "seg0", "seg1", "seg2", "ntc", "ntr", "lk"), list(out)))


def _item(B, H, W, idx, lagged=0):
    out = (ctypes.c_int * 10)()
    capi.load().cca_b200_decode_item(B, H, W, idx, lagged, out)
    return dict(zip(("col", "b", "line", "iq", "ik", "q0", "lq", "k0", "lk", "j"), list(out)))


@pytest.mark.parametrize("shape", SHAPES)
def test_items_cover_every_pair_exactly_once(shape):
    B, H, W = shape
    sp = _space(B, H, W)
    assert sp["total"] == B * sp["per_sample"] and sp["per_sample"] == sp["seg0"] + sp["seg1"] + sp["seg2"]
    assert sp["ntc"] == -(-H // 112) and sp["ntr"] == -(-W // 112)
    assert sp["lk"] in (80, 112)
    cover = {1: np.zeros((B, W, H, H), np.int16), 0: np.zeros((B, H, W, W), np.int16)} if H * W <= 200 * 200 else None
    prev_b = 0
    for idx in range(sp["total"]):
        it = _item(B, H, W, idx)
        L = H if it["col"] else W
        assert 0 < it["lq"] <= sp["lk"] and 0 < it["lk"] <= sp["lk"], it
        assert it["q0"] + it["lq"] <= L and it["k0"] + it["lk"] <= L
        assert it["j"] == idx - it["b"] * sp["per_sample"]
        assert it["b"] >= prev_b                                  # sample by sample
        prev_b = it["b"]
        # segment order inside a sample: column/first key block, other column items, row items
        seg = 0 if it["j"] < sp["seg0"] else (1 if it["j"] < sp["seg0"] + sp["seg1"] else 2)
        assert (seg == 0) == (it["col"] == 1 and it["ik"] == 0)
        assert (seg == 2) == (it["col"] == 0)
        if cover is not None:
            cover[it["col"]][it["b"], it["line"], it["q0"]:it["q0"] + it["lq"], it["k0"]:it["k0"] + it["lk"]] += 1
    if cover is not None:
        assert (cover[0] == 1).all() and (cover[1] == 1).all()


@pytest.mark.parametrize("shape", SHAPES)
def test_delta_producers_precede_their_consumers(shape):
    """Backward, delta_mode 1: the items that publish delta for a sample (column, first key block) all have a lower
    index than any other item of the sample, and every query pixel of the sample has exactly one producer."""
    B, H, W = shape
    sp = _space(B, H, W)
    seen = np.zeros((H, W), np.int32)
    for j in range(sp["per_sample"]):
        it = _item(B, H, W, j)
        if it["col"] and it["ik"] == 0:
            assert j < sp["seg0"]
            seen[it["q0"]:it["q0"] + it["lq"], it["line"]] += 1
        else:
            assert j >= sp["seg0"]
    assert (seen == 1).all()


def test_zero_shares_tile_the_sample():
    for (B, H, W), C, es in (((8, 97, 97), 512, 4), ((8, 97, 97), 64, 4), ((2, 5, 6), 64, 2), ((1, 193, 193), 512, 4)):
        sp = _space(B, H, W)
        sample = H * W * C * es
        share = -(-sample // sp["per_sample"])
        share = -(-share // 128) * 128                        # cca_items.cuh: zero_share_bytes
        covered = 0
        for j in range(sp["per_sample"]):
            lo, hi = j * share, min(sample, (j + 1) * share)
            if lo < hi:
                assert lo == covered and (hi - lo) % 16 == 0
                covered = hi
        assert covered == sample


@pytest.mark.parametrize("shape", [(1, 5, 6), (3, 5, 6), (8, 97, 97), (2, 129, 129), (4, 1, 300)])
def test_lagged_order_is_a_dependency_safe_permutation(shape):
    """decode_item_lagged: same items as the plain order; every producer (column, first key block) of a sample comes before
    every consumer of that sample (consumers only ever wait for lower indices), and the consumers of sample b come after
    the producers of sample b+1 (the one-block lag)."""
    B, H, W = shape
    sp = _space(B, H, W)
    key = lambda it: (it["b"], it["j"])
    plain = sorted(key(_item(B, H, W, i)) for i in range(sp["total"]))
    lag = [_item(B, H, W, i, 1) for i in range(sp["total"])]
    assert sorted(key(it) for it in lag) == plain
    last_prod, first_cons = {}, {}
    for i, it in enumerate(lag):
        if it["col"] and it["ik"] == 0:
            last_prod[it["b"]] = i
        else:
            first_cons.setdefault(it["b"], i)
    for b in range(B):
        assert last_prod[b] < first_cons[b]
        if b + 1 < B:
            assert last_prod[b + 1] < first_cons[b]
