"""CPU model of the statistics groups of cca_tc_stats.cu: ONE pass over a logit row in 16-column chunks (running max of the raw
logits, two exp2 accumulators rescaled when the max moves, masked entries as -inf) must equal the two-pass log-sum-exp the
oracle uses (cc_attention/functions.py:38-40 restated in oracle/cca_oracle.py), including rows whose first chunks are fully masked
and rows with no valid key at all."""
import math

import numpy as np
import pytest

LOG2E = 1.4426950408889634


def online_parts(s, valid):
    """float32 arithmetic in the kernel's order; returns log2 sum_j 2^(s_j log2e) over the valid entries (-inf if none)."""
    s = s.astype(np.float32)
    m = np.float32(-np.inf)
    l0 = l1 = np.float32(0.0)
    for c0 in range(0, len(s), 16):
        chunk = np.where(valid[c0:c0 + 16], s[c0:c0 + 16], np.float32(-np.inf)).astype(np.float32)
        chunk = np.pad(chunk, (0, 16 - len(chunk)), constant_values=-np.inf)
        cm = chunk.max()
        if cm > m:
            sc = np.float32(0.0) if m == -np.inf else np.exp2((m - cm) * np.float32(LOG2E), dtype=np.float32)
            l0, l1, m = l0 * sc, l1 * sc, cm
        nm = np.float32(0.0) if m == -np.inf else -m * np.float32(LOG2E)
        with np.errstate(over="ignore"):
            e = np.exp2(chunk * np.float32(LOG2E) + nm, dtype=np.float32)
        l0, l1 = l0 + e[0::2].sum(dtype=np.float32), l1 + e[1::2].sum(dtype=np.float32)
    l = l0 + l1
    return float(m * np.float32(LOG2E) + np.log2(l)) if l > 0 else -math.inf


@pytest.mark.parametrize("n,scale,seed", [(97, 3.0, 0), (112, 12.0, 1), (65, 30.0, 2), (17, 0.1, 3), (1, 5.0, 4)])
def test_single_pass_equals_two_pass_logsumexp(n, scale, seed):
    rng = np.random.default_rng(seed)
    s = (rng.standard_normal(n) * scale).astype(np.float32)
    for self_idx in (-1, 0, n // 2, n - 1):
        valid = np.ones(n, bool)
        if self_idx >= 0:
            valid[self_idx] = False
        ref = np.log2(np.sum(np.exp(s[valid].astype(np.float64)))) if valid.any() else -math.inf
        got = online_parts(s, valid)
        if ref == -math.inf:
            assert got == -math.inf
        else:
            assert abs(got - ref) <= 2e-5 * max(1.0, abs(ref)), (n, self_idx, got, ref)


def test_rows_that_start_with_masked_chunks_and_rows_without_keys():
    s = np.linspace(-40, 40, 64).astype(np.float32)
    valid = np.zeros(64, bool)
    valid[40:] = True                                  # the first two chunks contribute nothing: m stays -inf, l stays 0
    ref = np.log2(np.sum(np.exp(s[valid].astype(np.float64))))
    assert abs(online_parts(s, valid) - ref) <= 2e-5 * abs(ref)
    assert online_parts(s, np.zeros(64, bool)) == -math.inf
    # a late, much larger logit rescales everything before it to (almost) nothing without overflow
    s2 = np.full(48, -80.0, np.float32)
    s2[47] = 85.0
    assert abs(online_parts(s2, np.ones(48, bool)) - 85.0 * LOG2E) <= 1e-4
