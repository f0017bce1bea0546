"""The arithmetic claim behind gcnk_sequential_sum (csrc/elementwise.cu), checked on the CPU with numpy:

while the running fp32 sum S stays inside one binade, S is a multiple of u = ulp(S) and fl(S + t) = S + u * rint(t / u) for
every term t >= 0 that is not an exact tie — independent of S — so a block of terms can be rounded in parallel, summed as
integers and added to the mantissa of S; anything else (a tie, a negative or huge term, a carry out of the mantissa, S = 0)
is replayed with plain sequential additions.  The result must be bit-identical to the reference's scalar loop
(`total_loss += ...`, /root/reference/src/seq/module.cpp:125-143), i.e. to np.add.accumulate in float32.

This is a restatement of the kernel's block logic (block = 1,024 terms, fallback granularity 32), not the kernel itself:
the kernel is compared with the same scalar loop on the GPU in tests/test_gpu_ops.py::test_sequential_sum_bit_exact."""
import zlib

import numpy as np
import pytest


def scalar_loop(terms):
    s = np.float32(0.0)
    for t in terms.astype(np.float32):
        s = np.float32(s + t)
    return s


def fast_block(S, t):
    """One speculative block: returns the new S, or None when the block must be replayed."""
    sb = np.float32(S).view(np.uint32)
    e = int((sb >> 23) & 0xFF)
    if (sb >> 31) or e < 30 or e > 250:
        return None
    inv_u = np.uint32((127 + 150 - e) << 23).view(np.float32)            # 2^(150 - e) = 1 / ulp(S)
    x = t.astype(np.float32) * inv_u                                         # a power-of-two scaling: exact unless it over/underflows
    with np.errstate(invalid="ignore"):
        r = np.rint(x)
        if not (np.all(x >= 0) and np.all(x < 1048576.0)) or np.any(np.abs(x - r) == 0.5):
            return None
    R = int(r.astype(np.int64).sum())
    m = int(sb & 0x7FFFFF) | 0x800000
    if m + R >= 0x1000000:
        return None                                                          # would leave the binade
    return np.uint32((int(sb) & 0x7F800000) | ((m + R) & 0x7FFFFF)).view(np.float32)


def block_sum(terms, block=1024, row=32):
    S = np.float32(0.0)
    stats = {"fast": 0, "rows": 0, "scalar": 0}
    for b in range(0, len(terms), block):
        blk = terms[b:b + block]
        new = fast_block(S, blk)
        if new is not None:
            S = new
            stats["fast"] += 1
            continue
        for r0 in range(0, len(blk), row):
            rw = blk[r0:r0 + row]
            new = fast_block(S, rw)
            if new is not None:
                S = new
                stats["rows"] += 1
            else:
                for t in rw:
                    S = np.float32(S + np.float32(t))
                stats["scalar"] += 1
    return S, stats


@pytest.mark.parametrize("case", ["loss_like", "small_first", "ties", "with_negative", "with_huge", "denormals", "short"])
def test_block_parallel_sum_is_the_scalar_loop(case):
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    n = 40_000
    if case == "loss_like":                      # what the kernel sees in epoch 1: 153,756 terms, every one within 1e-3 of ln 41
        n = 153_756
        t = (np.log(41.0) + 1e-3 * rng.standard_normal(n)).astype(np.float32)
    elif case == "small_first":
        t = np.concatenate([1e-6 * rng.random(5000), 3.0 + rng.random(n - 5000)]).astype(np.float32)
    elif case == "ties":                         # multiples of 1/64: exact ties once ulp(S) reaches 1/32
        t = (rng.integers(1, 512, n) / 64.0).astype(np.float32)
    elif case == "with_negative":
        t = (3.0 + rng.random(n)).astype(np.float32)
        t[rng.integers(0, n, 20)] *= -1
    elif case == "with_huge":
        t = (3.0 + rng.random(n)).astype(np.float32)
        t[[100, 20_000]] = [1e30, 7e5]
    elif case == "denormals":
        t = np.concatenate([np.full(3000, 1e-42), 2.0 + rng.random(n - 3000)]).astype(np.float32)
    else:
        t = (3.0 + rng.random(77)).astype(np.float32)
    want = scalar_loop(t)
    got, stats = block_sum(t)
    assert np.float32(got).view(np.uint32) == np.float32(want).view(np.uint32), (case, got, want, stats)
    if case == "loss_like":
        assert stats["fast"] >= 130             # the fast path carries the sum: only binade crossings fall back
        # and the scalar loop really is further from the exact sum than the 1e-4 parity tolerance: a more accurate
        # (tree) sum would NOT match gcn-seq here
        exact = float(np.sum(t.astype(np.float64)))
        assert abs(float(want) - exact) / exact > 1e-4
