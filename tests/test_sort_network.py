"""The index arithmetic of the detection tail's bitonic network (csrc/rn_decode.cu, bitonic_sort_u64), emulated in
NumPy -- CPU only.  The kernel runs the strides below 32 in registers (partner = lane ^ j, keep the minimum when
`((lane & j) == 0) == ((i & k) == 0)`) and the strides >= 32 in shared memory with one pair per thread
(i = t with a zero bit inserted at log2(j), partner i | j, swap when `(a > b) == ((i & k) == 0)`); this checks that
the two forms together are a sorting network for every power-of-two size the kernel uses."""
import numpy as np
import pytest


def _register_stage(key, k, j):
    i = np.arange(key.shape[0])
    other = key[i ^ j]
    up = (i & k) == 0
    take_min = (((i % 32) & j) == 0) == up
    return np.where(take_min, np.minimum(key, other), np.maximum(key, other))


def _shared_stage(key, k, j):
    t = np.arange(key.shape[0] // 2)
    i = ((t & ~(j - 1)) << 1) | (t & (j - 1))
    a, b = key[i], key[i | j]
    swap = (a > b) == ((i & k) == 0)
    out = key.copy()
    out[i] = np.where(swap, b, a)
    out[i | j] = np.where(swap, a, b)
    return out


def _network(key):
    n = key.shape[0]
    k = 2
    while k <= 32:                      # first pass: every aligned run of 32 keys, all in registers
        j = k >> 1
        while j > 0:
            key = _register_stage(key, k, j)
            j >>= 1
        k <<= 1
    k = 64
    while k <= n:
        j = k >> 1
        while j >= 32:
            key = _shared_stage(key, k, j)
            j >>= 1
        while j > 0:
            key = _register_stage(key, k, j)
            j >>= 1
        k <<= 1
    return key


@pytest.mark.parametrize('n_pad', [32, 64, 128, 1024, 2048, 8192])
def test_bitonic_network_sorts(n_pad):
    rng = np.random.default_rng(n_pad)
    for fill in (n_pad, n_pad - 1, n_pad // 2 + 3, 1, 0):
        key = np.full(n_pad, np.iinfo(np.uint64).max, np.uint64)          # padding = ~0
        # distinct keys (the kernel's keys carry a unique index in their low bits)
        vals = rng.choice(1 << 40, size=fill, replace=False).astype(np.uint64) if fill else np.zeros(0, np.uint64)
        key[:fill] = vals
        rng.shuffle(key)
        out = _network(key)
        assert np.array_equal(out, np.sort(key)), (n_pad, fill)
