"""CPU checks of the arithmetic identities behind ``ties_one_fast`` (modelcompose_b200/csrc/mc_ties_kernels.cuh).

The GPU kernel replaces the compare / select / min-max formulation of the TIES element (oracle/ties_oracle.py,
reference ties_merging.py:98-155) by FMUL / FADD / FFMA forms.  Two things are proven here without a GPU:

* a numpy float32 emulation of the kernel's exact instruction sequence (every op one IEEE round-to-nearest step, FMA
  emulated where the product is exact) reproduces the oracle bit for bit on random, cancelling, subnormal and
  overflowing 16-bit inputs, for both default signs;
* the division-free MEAN quotient ``q1 = fma(fma(-q0, c, x), r, q0)``, ``q0 = x * r`` is the correctly rounded
  ``x / c`` for EVERY non-negative finite bf16 / fp16 value x, every count c <= 8 and every reciprocal r within 2 ulp
  of 1 / c (the hardware's rcp.approx is within 1 ulp) — exhaustive.
"""
import numpy as np
import pytest
import torch

from oracle import ties_oracle as TO

F32 = np.float32
INF = F32(np.inf)


def mul_sat(a, b):
    """PTX mul.rn.sat.f32: clamp to [0, 1], NaN -> +0."""
    with np.errstate(invalid="ignore", over="ignore"):
        r = (a.astype(F32) * F32(b)).astype(F32)
    r = np.where(np.isnan(r), F32(0), r)
    return np.clip(r, F32(0), F32(1)).astype(F32)


def round_dt(x, dt):
    return torch.from_numpy(x.astype(F32)).to(dt).to(torch.float32).numpy()


def rn32_sum(hi_terms):
    """RN32(a + b) for float64 a, b whose exact sum may need more than 53 bits: TwoSum + midpoint tie-break."""
    a, b = hi_terms
    s = a + b
    bb = s - a
    err = (a - (s - bb)) + (b - bb)          # exact error of the float64 addition
    c32 = s.astype(F32)
    d = s - c32.astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        other = np.nextafter(c32, np.where(d > 0, INF, -INF).astype(F32)).astype(np.float64)
    mid = (c32.astype(np.float64) + other) / 2
    is_mid = (d != 0) & (s == mid)
    toward_other = is_mid & (np.sign(err) == np.sign(d)) & (err != 0)
    # at an exact float32 midpoint the float64 rounding error decides; numpy's own tie-to-even is right when err == 0
    away_c32 = is_mid & (np.sign(err) == -np.sign(d)) & (err != 0)
    out = np.where(toward_other, other.astype(F32), c32)
    out = np.where(away_c32, c32, out)
    return out.astype(F32)


def fast_divide(x, c, r):
    """The kernel's MEAN quotient (x >= 0 float32 holding a 16-bit value, c float32 count, r float32 ~ 1 / c)."""
    x64, c64, r64 = x.astype(np.float64), c.astype(np.float64), r.astype(np.float64)
    q0 = (x64 * r64).astype(F32)                           # exact product (<= 11 + 24 bits), one rounding
    res = (x64 - q0.astype(np.float64) * c64).astype(F32)  # fma(-q0, c, x): product exact, difference exact in float64
    q1 = rn32_sum((q0.astype(np.float64), res.astype(np.float64) * r64))  # fma(res, r, q0): product exact (<= 48 bits)
    return np.fmin(q1, x)                                  # fminf: returns x when q1 is NaN (x = inf)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_mean_quotient_is_correctly_rounded_exhaustive(dt):
    bits = torch.arange(0, 1 << 15, dtype=torch.int32).to(torch.int16)
    x = bits.view(dt).to(torch.float32).numpy()
    x = x[np.isfinite(x)]
    assert x.size > 30000 and x.min() == 0.0
    for c in range(1, 9):
        cf = np.full_like(x, c, dtype=F32)
        want = (x / cf).astype(F32)                        # IEEE float32 division
        r_exact = F32(1.0) / F32(c)
        for ulps in (-2, -1, 0, 1, 2):
            r = r_exact
            for _ in range(abs(ulps)):
                r = np.nextafter(r, INF if ulps > 0 else -INF, dtype=F32)
            got = fast_divide(x, cf, np.full_like(x, r))
            bad = got.view(np.uint32) != want.view(np.uint32)
            assert not bad.any(), (dt, c, ulps, x[bad][:4], got[bad][:4], want[bad][:4])
    # overflowed 16-bit sum: x = inf stays inf
    got = fast_divide(np.array([np.inf], F32), np.array([2], F32), np.array([0.5], F32))
    assert np.isinf(got[0]) and got[0] > 0


def emulate_fast(flat: torch.Tensor, thr, majority: float, func: str, packed_trim: bool = False):
    """ties_one_fast, op for op, on [n_src, d] 16-bit inputs; returns (out, p, n, amb).  ``packed_trim``: the vector path's
    trim on the packed words (``w & mask(|x| >= thr)``: trimmed entries become +0 instead of x * 0 = +-0)."""
    dt = flat.dtype
    x = flat.to(torch.float32).numpy()
    n_src = x.shape[0]
    mh = F32(0.5 if majority > 0 else -0.5)
    with np.errstate(invalid="ignore", over="ignore"):
        if packed_trim:
            m = [np.where(np.abs(x[s]) >= F32(thr[s]), x[s], F32(0)).astype(F32) for s in range(n_src)]
        else:
            m = [(x[s] * (np.abs(x[s]) >= F32(thr[s])).astype(F32)).astype(F32) for s in range(n_src)]
        acc = m[0]
        for s in range(1, n_src):
            acc = (acc + m[s]).astype(F32)
        p, n = mul_sat(acc, INF), mul_sat(acc, -INF)
        hs = (-n * (F32(0.5) + mh) + (p * (F32(0.5) - mh) + mh)).astype(F32)   # every step exact
        if majority > 0:   # the speculative pass (default sign +) computes it as 0.5 - n
            assert np.array_equal(hs, (F32(0.5) - n).astype(F32))
        sg = (hs + hs).astype(F32)
        ksum = cnt = None
        for s in range(n_src):
            k = (F32(0.5) * np.abs(m[s]) + (m[s] * hs).astype(F32)).astype(F32)  # both products exact: one rounding
            ksum = k if s == 0 else (np.maximum(ksum, k) if func == "max" else (ksum + k).astype(F32))
            one = mul_sat(k, INF)
            cnt = one if s == 0 else (cnt + one).astype(F32)
        some = mul_sat(cnt, 1.0) if func == "mean" else mul_sat(ksum, INF)
        amb = (-some * (p + n) + some).astype(F32)
        if func == "sum":
            out = torch.from_numpy((ksum * sg + F32(0)).astype(F32)).to(dt)
        elif func == "max":
            o = (ksum * sg).astype(F32)
            if majority == 0:   # ties_one: the reference multiplies by sign 0 where no sign was elected
                o = np.where((p == 0) & (n == 0), F32(0), o).astype(F32)
            out = torch.from_numpy(o).to(dt)
        else:
            xr = round_dt(ksum, dt)
            assert not np.any((cnt == 0) & (xr != 0))       # nothing kept: the kept sum is +0 ...
            with np.errstate(divide="ignore"):
                q = np.fmin((xr / cnt).astype(F32), xr)     # ... 0 / 0 = NaN and fmin(NaN, 0) = 0: no max(cnt, 1) needed.  The quotient itself
                                                            # is covered by the exhaustive test
            out = torch.from_numpy((q * sg + F32(0)).astype(F32))
    return out, p, n, amb


def make_flat(kind, n_src, d, dt, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "gauss":
        t = torch.randn(n_src, d, generator=g) * 0.02
    elif kind == "ints":
        t = torch.randint(-3, 4, (n_src, d), generator=g).float()
    elif kind == "neg":
        t = torch.randn(n_src, d, generator=g) * 0.02 - 0.03
    elif kind == "subnormal":
        tiny = 2.0 ** -130 if dt == torch.bfloat16 else 2.0 ** -22
        t = torch.randint(-6, 7, (n_src, d), generator=g).float() * tiny
    else:  # overflowing sums in fp16, wide dynamic range in bf16
        t = (torch.randn(n_src, d, generator=g) * torch.exp(torch.randn(n_src, d, generator=g) * 6.0)).clamp(-6e4, 6e4)
    return t.to(dt)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("kind,K", [("gauss", 20), ("ints", 50), ("ints", 99), ("neg", 20), ("subnormal", 60), ("wide", 30)])
@pytest.mark.parametrize("n_src", [1, 2, 3, 5, 8])
def test_fast_formulation_matches_oracle(dt, kind, K, n_src):
    flat = make_flat(kind, n_src, 20011, dt, seed=31 * n_src + len(kind))
    for func in ("sum", "mean", "max"):
        want, st = TO.ties_merge_flat(flat, K, func)
        thr = st["thresholds"].numpy()
        for majority in (st["majority"], -st["majority"] if st["majority"] else 1.0, 0.0):
            ref = TO.merge_given_statistics(flat, thr, majority, func)
            for packed in (False, True):   # scalar tail path (multiply) and vector path (packed mask) of the kernel
                got, p, n, amb = emulate_fast(flat, thr, majority, func, packed_trim=packed)
                iv = torch.int16 if got.dtype != torch.float32 else torch.int32
                assert got.dtype == ref.dtype
                assert torch.equal(got.view(iv), ref.view(iv)), (func, majority, packed, int((got.float() != ref.float()).sum()))
        got, p, n, amb = emulate_fast(flat, thr, st["majority"], func)
        assert (int(p.sum()), int(n.sum()), int(amb.sum())) == (st["n_pos"], st["n_neg"], st["ambiguous"])
        assert set(np.unique(np.concatenate([p, n, amb]))) <= {0.0, 1.0}


def test_packed_threshold_counter_has_no_cross_half_borrow():
    """ties_count_kernel counts ">= t" for two 15-bit keys per 32-bit word: k = word | 0x80008000, flags = (k - (t | t << 16))
    & 0x80008000, summed with dp4a (0x80 per flag).  Every low key x a spread of thresholds (incl. 0, 1, 0x7fff and the clamp
    0x8000) x high keys at the extremes: the flag of each half equals key >= t, i.e. no borrow ever crosses the halves."""
    lo_keys = np.arange(1 << 15, dtype=np.uint32)
    rng = np.random.default_rng(3)
    thresholds = np.unique(np.concatenate([[0, 1, 2, 0x3f80, 0x7ffe, 0x7fff, 0x8000], rng.integers(0, 0x8001, 200)])).astype(np.uint32)
    hi_keys = np.array([0, 1, 0x3f80, 0x7ffe, 0x7fff], dtype=np.uint32)
    for sign_bits in (0x00000000, 0x80008000, 0x80000000):       # the sign bits of the data are overwritten by the OR
        for hk in hi_keys:
            words = (lo_keys | (hk << 16) | np.uint32(sign_bits)).astype(np.uint32)
            k = words | np.uint32(0x80008000)
            for t in thresholds:
                t2 = np.uint32(t | (t << 16))
                flags = (k - t2) & np.uint32(0x80008000)          # uint32 wrap-around arithmetic, as the GPU
                assert np.array_equal((flags & 0x8000) != 0, lo_keys >= t), (hk, t)
                assert np.all(((flags >> 31) & 1) == (1 if hk >= t else 0)), (hk, t)
                # dp4a with 0x01010101 adds the four bytes: 0x80 per set flag
                dp = (flags & 0xff) + ((flags >> 8) & 0xff) + ((flags >> 16) & 0xff) + ((flags >> 24) & 0xff)
                assert np.array_equal(dp >> 7, (lo_keys >= t).astype(np.uint32) + (1 if hk >= t else 0))


def test_bf16_trim_as_fp16_bit_patterns():
    """trim_word compares bf16 pairs with the fp16 form of HSET2 (the one whose |.| operand modifier ptxas folds): for every magnitude
    pattern and every threshold pattern below 0x7c00, "not (|x| < t) as fp16" equals "|x| >= t as bf16"; patterns from 0x7c01 up are
    NaN to the fp16 compare and come out as kept, which is what bf16 says for finite values beyond 2^121 and for inf."""
    mags = np.arange(0, 0x8000, dtype=np.uint16)
    as_bf16 = (mags.astype(np.uint32) << 16).view(np.float32)
    as_fp16 = mags.view(np.float16)
    rng = np.random.default_rng(0)
    thr = np.unique(np.concatenate([rng.integers(0, 0x7c00, 600), [0, 1, 0x3ff, 0x400, 0x7bff]]).astype(np.uint16))
    with np.errstate(invalid="ignore"):
        for t in thr:
            t_bf16 = np.uint32(int(t) << 16).view(np.float32)
            keep_ref = as_bf16 >= t_bf16                      # NaN patterns (> 0x7f80): False in the reference, see below
            keep_dev = ~(as_fp16 < np.uint16(t).view(np.float16))
            finite_or_inf = mags <= 0x7f80
            assert np.array_equal(keep_ref[finite_or_inf], keep_dev[finite_or_inf]), hex(int(t))
            assert keep_dev[~finite_or_inf].all()             # bf16 NaNs stay in the sum (the reference's x * mask keeps them NaN too)


def test_census_bit_masks_in_float_accumulators():
    """ties_chunk_fast keeps the census of 16 elements as bit masks built by FFMAs (mask += indicator * 2^bit): the sums stay below
    2^16, exact in fp32 whatever the order; counts are popcounts and the class-3 set is some & ~(p | n)."""
    rng = np.random.default_rng(3)
    for _ in range(2000):
        acc_sign = rng.integers(-1, 2, 16)                     # sign of the trimmed sum per element
        kept = rng.integers(0, 2, 16) | (acc_sign != 0)        # an elected sign implies a kept entry
        p, n, some = (acc_sign > 0).astype(F32), (acc_sign < 0).astype(F32), kept.astype(F32)
        m_pos = m_neg = m_some = F32(0)
        for b in rng.permutation(16):
            bit = F32(1 << int(b))
            m_pos, m_neg, m_some = F32(p[b] * bit + m_pos), F32(n[b] * bit + m_neg), F32(some[b] * bit + m_some)
        bp, bn, bs = int(m_pos), int(m_neg), int(m_some)
        assert bin(bp).count("1") == int(p.sum()) and bin(bn).count("1") == int(n.sum())
        amb = bs & ~(bp | bn)
        assert [b for b in range(16) if amb >> b & 1] == [b for b in range(16) if kept[b] and acc_sign[b] == 0]
