"""GPU parity of the N-source merge kernel (through the C ABI) against oracle/merge_oracle.py.

Bar: bit-exact (integer compare of the raw words; NaN positions compared as NaN-ness)."""

import pytest
import torch

from modelcompose_b200 import merge as M
from oracle import merge_oracle as MO

pytestmark = pytest.mark.gpu

INT_VIEW = {torch.float32: torch.int32, torch.float16: torch.int16, torch.bfloat16: torch.int16}
SIZES = [0, 1, 7, 15, 16, 17, 4096, 4097, 65536 + 3, (1 << 20) + 5]


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    a, b = a.cpu(), b.cpu()
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    if not torch.equal(nan_a, nan_b):
        return False
    av, bv = a.view(INT_VIEW[a.dtype]), b.view(INT_VIEW[b.dtype])
    return torch.equal(av[~nan_a], bv[~nan_b])


def make_sources(n_src, sizes, dtype, seed, special=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for s in range(n_src):
        lst = []
        for n in sizes:
            t = torch.randn(n, generator=g) * 0.02
            if special and n >= 16:
                t[:8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1e-40, -1e-40, 3.0e38, 65504.0])
                t[8:12] = torch.tensor([1.0, -1.0, 0.333, 2.0 ** -14]) * (s + 1)
            lst.append(t.to(dtype))
        out.append(lst)
    return out


@pytest.mark.parametrize("src_dtype,dst_dtype", [
    (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16), (torch.float32, torch.float32),
    (torch.bfloat16, torch.float32), (torch.float16, torch.float32), (torch.float32, torch.bfloat16),
    (torch.float32, torch.float16)])
@pytest.mark.parametrize("n_src", [1, 2, 3, 4, 5, 8])
def test_weighted_bit_exact(src_dtype, dst_dtype, n_src):
    srcs = make_sources(n_src, SIZES, src_dtype, seed=n_src)
    weights = [0.333, 0.333, 0.333, 0.001, -1.5, 0.25, 1.0, 2.0][:n_src]
    dev = [[t.cuda() for t in lst] for lst in srcs]
    outs = [torch.full((n,), 7.0, dtype=dst_dtype, device="cuda") for n in SIZES]
    plan = M.MergePlan(dev, outs)
    plan.run(weights, "weighted")
    torch.cuda.synchronize()
    assert plan.algorithmic_bytes == sum(SIZES) * (n_src * srcs[0][0].element_size() + outs[0].element_size())
    for t, n in enumerate(SIZES):
        want = MO.weighted_merge([srcs[s][t] for s in range(n_src)], weights, dst_dtype)
        assert bits_equal(outs[t], want), (n, src_dtype, dst_dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n_src", [1, 2, 3, 4])
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_reference_sum_mean_bit_exact(dtype, n_src, mode):
    srcs = make_sources(n_src, SIZES, dtype, seed=10 + n_src)
    dev = [[t.cuda() for t in lst] for lst in srcs]
    outs = [torch.empty(n, dtype=dtype, device="cuda") for n in SIZES]
    M.MergePlan(dev, outs).run(None, mode)
    torch.cuda.synchronize()
    for t, n in enumerate(SIZES):
        want = MO.ref_sum([srcs[s][t] for s in range(n_src)])
        if mode == "mean":
            want = want / n_src
        if not torch.is_tensor(want):  # n == 0 never reaches here; n_src >= 1
            want = torch.as_tensor(want)
        assert bits_equal(outs[t], want), (n, dtype, mode)


X = 1 << 24  # explicit tuning code


@pytest.mark.parametrize("tuning", [0, X | 0, X | 1, X | 2, X | 3, X | 4, X | 5, X | (1 << 16), X | 1 | (2 << 8)])
def test_every_tuning_variant(tuning):
    sizes = [(1 << 21) + 77, 12345, 4096 * 11008 // 64]
    srcs = make_sources(3, sizes, torch.bfloat16, seed=99, special=False)
    dev = [[t.cuda() for t in lst] for lst in srcs]
    outs = [torch.empty(n, dtype=torch.bfloat16, device="cuda") for n in sizes]
    w = [0.333, 0.333, 0.333]
    M.MergePlan(dev, outs, tuning=tuning).run(w)
    torch.cuda.synchronize()
    for t in range(len(sizes)):
        assert bits_equal(outs[t], MO.weighted_merge([srcs[s][t] for s in range(3)], w))


def test_unaligned_views_and_fused_arena():
    # tensors carved back-to-back out of one arena per source (fused into one segment) and views that start
    # at odd element offsets (scalar path)
    sizes = [4096, 33, 8192, 1]
    g = torch.Generator().manual_seed(5)
    total = sum(sizes)
    arenas = [(torch.randn(total + 3, generator=g) * 0.02).to(torch.bfloat16) for _ in range(3)]
    w = [0.5, 0.25, 0.25]
    for shift in (0, 1, 3):
        dev_arenas = [a.cuda() for a in arenas]
        out_arena = torch.zeros(total + 3, dtype=torch.bfloat16, device="cuda")
        offs = [shift + sum(sizes[:i]) for i in range(len(sizes))]
        dev = [[a[o:o + n] for o, n in zip(offs, sizes)] for a in dev_arenas]
        outs = [out_arena[o:o + n] for o, n in zip(offs, sizes)]
        M.MergePlan(dev, outs).run(w)
        torch.cuda.synchronize()
        want = MO.weighted_merge([a[shift:shift + total] for a in arenas], w)
        assert bits_equal(out_arena[shift:shift + total], want)
        assert out_arena[:shift].abs().sum().item() == 0 and out_arena[shift + total:].abs().sum().item() == 0


def test_host_streaming_merge_matches_device():
    sizes = [0, 5, 100_000, 3_000_000, 17]
    srcs = make_sources(3, sizes, torch.bfloat16, seed=3, special=False)
    w = [0.333, 0.333, 0.333]
    outs = M.merge_host_tensors(srcs, w, "weighted", staging_bytes=1 << 20)  # small slabs → many pipeline steps
    for t in range(len(sizes)):
        assert bits_equal(outs[t], MO.weighted_merge([srcs[s][t] for s in range(3)], w))
    outs = M.merge_host_tensors(srcs, None, "mean", staging_bytes=1 << 20)
    for t in range(len(sizes)):
        if sizes[t]:
            assert bits_equal(outs[t], MO.ref_sum([srcs[s][t] for s in range(3)]) / 3)


def test_merge_state_dicts_device_materialises_reset_blend():
    # W_eff = (1-Σw)·W + Σ w_m·ckpt_m with ckpt_m = W + (α/r)·B_m A_m equals W + Σ w_m (α/r) B_m A_m
    g = torch.Generator().manual_seed(0)
    W = torch.randn(64, 48, generator=g) * 0.02
    deltas = [2.0 * (torch.randn(64, 8, generator=g) * 0.02) @ (torch.rand(8, 48, generator=g) - 0.5) for _ in range(3)]
    ws = [0.333, 0.333, 0.333]
    sds = [{"w": W.cuda()}] + [{"w": (W + d).cuda()} for d in deltas]
    weights = [1.0 - sum(ws)] + ws
    out = M.merge_state_dicts_device(sds, weights)["w"].cpu()
    want = MO.weighted_merge([W] + [W + d for d in deltas], weights)
    assert bits_equal(out, want)
    direct = W + sum(w * d for w, d in zip(ws, deltas))
    assert (out - direct).abs().max().item() < 1e-6


def test_large_linearity_property():
    # size-independent property at a 7B-layer-sized tensor: merge(a,b,c; w) with w=(1,0,0) returns a exactly,
    # and merging [x, x, x] with weights summing to 1 in exact binary fractions returns x.
    n = 11008 * 4096
    x = (torch.randn(n, device="cuda") * 0.02).to(torch.bfloat16)
    y = (torch.randn(n, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.empty_like(x)
    M.MergePlan([[x], [y], [y]], [out]).run([1.0, 0.0, 0.0])
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), (x + 0.0).view(torch.int16)) or torch.equal(out, x)
    M.MergePlan([[x], [x], [x]], [out]).run([0.5, 0.25, 0.25])
    torch.cuda.synchronize()
    assert torch.equal(out, x)


def test_rejects_cpu_tensors_and_bad_modes():
    a = torch.zeros(4, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        M.MergePlan([[a]], [a.clone()])
    x = torch.zeros(4, dtype=torch.bfloat16, device="cuda")
    o = torch.zeros(4, dtype=torch.float32, device="cuda")
    with pytest.raises(Exception, match="src dtype == dst dtype"):
        M.MergePlan([[x]], [o]).run(None, "sum")
