"""GPU parity of the tcgen05 routed/grouped linear kernels and the row-wise glue (through the C ABI).

Floating point: the kernels accumulate in fp32 and round ONCE to bf16/fp16, the reference's eager path rounds after
every op.  Tolerance (stated per north_star): against an fp32 evaluation of the same formula on the same 16-bit
inputs, |got - ref| <= 2^-7*|ref| + 2^-8*max|ref| (bf16; = 1 ulp of the element plus half an ulp of the row scale),
fp16: 2^-10*|ref| + 2^-11*max|ref|.  Against the reference's own 16-bit outputs (golden fixtures) the bound is
2^-6 (bf16) / 2^-9 (fp16) of the output scale, the same bar tests/test_oracle_golden.py uses for the oracle."""
import pytest
import torch

from modelcompose_b200 import _cabi
from modelcompose_b200 import linear as LN
from oracle import model_oracle as XO

pytestmark = pytest.mark.gpu

RTOL = {torch.bfloat16: 2 ** -7, torch.float16: 2 ** -10}


def check(got, ref, dtype, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs()
    bound = RTOL[dtype] * ref.abs() + RTOL[dtype] / 2 * scale
    bad = err > bound
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance, max err {err.max().item():.4g} at scale {scale:.4g}"


def rnd(shape, dtype, seed, std=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * std).to(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,tuning", [(1, 8, 8, 0), (128, 256, 64, 0), (129, 264, 72, 0), (48, 192, 256, 0), (300, 136, 688, 1),
                                          (1000, 1000, 1000, 2), (513, 4096, 384, 0), (960, 688, 256, 0), (77, 32000, 256, 0),
                                          # CTA-pair kernel (cta_group::2)
                                          (1, 256, 64, 3), (300, 264, 72, 3), (1000, 1000, 1000, 3), (2100, 4096, 512, 3),
                                          # CTA-pair kernel, 256 x 256 pair tiles, overlapped epilogue
                                          (1, 256, 64, 4), (300, 264, 72, 4), (1000, 1000, 1000, 4), (2100, 4096, 512, 4), (40000, 512, 128, 4)])
def test_plain_linear(dtype, M, N, K, tuning):
    A, B = rnd((M, K), dtype, 1), rnd((N, K), dtype, 2, std=0.05)
    C = torch.full((M, N), 9.0, dtype=dtype, device="cuda")
    LN.LinearPlan([LN.Problem(A.cuda(), B.cuda(), C)], tuning=tuning).run()
    torch.cuda.synchronize()
    check(C, A.float() @ B.float().t(), dtype, f"{M}x{N}x{K}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_strided_views_and_epilogues(dtype):
    M, N, K = 200, 320, 136
    Abig, B = rnd((M, K + 24), dtype, 3).cuda(), rnd((N, K), dtype, 4, std=0.1).cuda()
    A = Abig[:, 8:8 + K]  # leading dimension > K, 16-byte aligned start
    bias = rnd((N,), dtype, 5).cuda()
    res = rnd((M, N), dtype, 6).cuda()
    ref = A.float() @ B.float().t()
    C = torch.empty((M, N + 8), dtype=dtype, device="cuda")[:, :N]
    LN.LinearPlan([LN.Problem(A, B, C, bias=bias, epilogue=LN.EPI_BIAS)]).run()
    check(C, ref + bias.float(), dtype, "bias")
    LN.LinearPlan([LN.Problem(A, B, C, bias=bias, epilogue=LN.EPI_BIAS_GELU)]).run()
    check(C, torch.nn.functional.gelu(ref + bias.float()), dtype, "bias+gelu")
    LN.LinearPlan([LN.Problem(A, B, C, residual=res, epilogue=LN.EPI_RESIDUAL)]).run()
    check(C, ref + res.float(), dtype, "residual")
    x = res.clone()  # in place: C aliases the residual (how the decoder layer uses it)
    LN.LinearPlan([LN.Problem(A, B, x, residual=x, epilogue=LN.EPI_RESIDUAL)]).run()
    check(x, ref + res.float(), dtype, "residual in place")


def test_compact_schedule_multi_problem():
    """Three ROWMASK problems in one launch (q/k/v down-projections) under the compacted tile schedule."""
    dtype, T, in_f, r = torch.bfloat16, 3000, 256, 128
    modal_names = ["default", "audio", "vision", "video"]
    dnames = [f"default-{m}" for m in modal_names[1:]]
    mid = make_modal_id(T, 11, 4)
    x = rnd((T, in_f), dtype, 1).cuda()
    rg = mid.cuda()
    mt = LN.route_tile_masks(rg)
    probs, refs = [], []
    for j in range(3):
        A = {n: rnd((r, in_f), dtype, 100 * j + i, 0.1) for i, n in enumerate(modal_names[1:] + dnames)}
        Bm = {n: rnd((64, r), dtype, 100 * j + 50 + i, 0.1) for i, n in enumerate(A)}
        scaling = {n: 1.0 + 0.25 * i for i, n in enumerate(A)}
        pk = LN.pack_adapters({k: v.cuda() for k, v in A.items()}, {k: v.cuda() for k, v in Bm.items()}, scaling, modal_names,
                              dnames, in_f, 64, dtype, torch.device("cuda"))
        Tb = torch.full((T, pk.A_all.shape[0]), float("nan"), dtype=dtype, device="cuda")
        probs.append(LN.Problem(x, pk.A_all, Tb, col_scale=pk.col_scale, row_group=rg, mtile_mask=mt, group_cols=pk.group_cols,
                                epilogue=LN.EPI_ROWMASK))
        full = (x.float() @ pk.A_all.float().t()) * pk.col_scale[None]
        colg = torch.zeros(pk.A_all.shape[0], dtype=torch.long)
        for g in range(4):
            colg[pk.group_cols[g]:pk.group_cols[g + 1]] = g
        refs.append((full.cpu() * (colg[None] == mid.long()[:, None]), colg))
    LN.LinearPlan(probs, tuning=1).run()
    torch.cuda.synchronize()
    for pr, (ref, colg) in zip(probs, refs):
        got = pr.C.float().cpu()
        own = colg[None] == mid.long()[:, None]
        check(torch.where(own, got, torch.zeros_like(got)), ref, dtype, "compact multi-problem")
        # columns of a group present in the row's 128-row tile but not the row's own are written as exact zeros
        tile_groups = torch.stack([torch.stack([(mid[t * 128:(t + 1) * 128] == g).any() for g in range(4)]) for t in range((T + 127) // 128)])
        present = tile_groups[torch.arange(T) // 128][:, colg]
        assert (got[present & ~own] == 0).all()


def test_silu_mul_epilogue_matches_separate_ops():
    dtype = torch.bfloat16
    M, N, K = 300, 688, 256
    x, Wg, Wu = rnd((M, K), dtype, 1).cuda(), rnd((N, K), dtype, 2, 0.2).cuda(), rnd((N, K), dtype, 3, 0.2).cuda()
    gate, up = torch.empty((M, N), dtype=dtype, device="cuda"), torch.empty((M, N), dtype=dtype, device="cuda")
    LN.LinearPlan([LN.Problem(x, Wg, gate), LN.Problem(x, Wu, up)]).run()
    want = LN.silu_mul(gate, up)
    act = gate.clone()
    LN.LinearPlan([LN.Problem(x, Wu, act, residual=act, epilogue=LN.EPI_SILU_MUL)]).run()  # in place over the gate buffer
    torch.cuda.synchronize()
    assert torch.equal(act, want)
    ref = torch.nn.functional.silu(gate.cpu()) * up.cpu()
    d = (act.cpu().float() - ref.float()).abs()
    assert (d <= 2 ** -6 * ref.float().abs() + 1e-7).all() and (d > 0).float().mean().item() < 0.01


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("tuning", [0, 1, 3, 4])
@pytest.mark.parametrize("M,N,K", [(300, 688, 256), (1030, 520, 128), (77, 264, 72), (512, 256, 64)])
def test_staged_epilogue_is_bit_identical_to_row_per_thread(dtype, tuning, M, N, K):
    """The coalesced (shared-memory transposed) NONE / RESIDUAL / SILU_MUL epilogues against the row-per-thread form
    (tuning bit 17) on every kernel, ragged M / N tails, scattered output rows."""
    x, W = rnd((M, K), dtype, 11).cuda(), rnd((N, K), dtype, 12, 0.2).cuda()
    res = rnd((M, N), dtype, 13).cuda()
    perm = torch.randperm(M, generator=torch.Generator().manual_seed(14)).to(torch.int32).cuda()
    for epi, rowmap in ((LN.EPI_NONE, None), (LN.EPI_RESIDUAL, None), (LN.EPI_SILU_MUL, None), (LN.EPI_NONE, perm)):
        outs = []
        for bit in (0, 1 << 17):
            out = res.clone() if epi != LN.EPI_NONE else torch.full((M, N), 7.0, dtype=dtype, device="cuda")
            LN.LinearPlan([LN.Problem(x, W, out, residual=out if epi != LN.EPI_NONE else None, epilogue=epi, c_rowmap=rowmap)],
                          tuning=tuning | bit).run()
            outs.append(out)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), (epi, tuning, rowmap is not None)
    ref = x.float() @ W.float().t()
    plain = torch.empty((M, N), dtype=dtype, device="cuda")
    LN.LinearPlan([LN.Problem(x, W, plain)], tuning=tuning).run()
    check(plain, ref, dtype, "staged plain epilogue")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("D,tuning", [(128, 0), (64, 0), (64, 1), (128, 3), (128, 4), (64, 4)])
def test_rope_epilogue_bit_exact_vs_separate_kernel(dtype, D, tuning):
    """q/k projection with RoPE in the epilogue == projection followed by mc_rope == oracle apply_rope on the rounded projection."""
    Bn, S, nH, K = 3, 50, 4, 192
    H = nH * D
    x, W = rnd((Bn * S, K), dtype, 1).cuda(), rnd((H, K), dtype, 2, 0.1).cuda()
    cos, sin = XO.rope_cos_sin(D, 128, dtype)
    cos_d, sin_d = cos.cuda(), sin.cuda()
    pos = torch.tensor([7], dtype=torch.int32, device="cuda")
    plain = torch.empty((Bn * S, H), dtype=dtype, device="cuda")
    LN.LinearPlan([LN.Problem(x, W, plain)], tuning=tuning).run()
    fused = torch.empty_like(plain)
    LN.LinearPlan([LN.Problem(x, W, fused, epilogue=LN.EPI_ROPE, rope=(cos_d, sin_d, pos, S, D))], tuning=tuning).run()
    torch.cuda.synchronize()
    q = plain.cpu().view(Bn, S, nH, D).transpose(1, 2)
    position_ids = (7 + torch.arange(S))[None].expand(Bn, S)
    want, _ = XO.apply_rope(q, q, cos, sin, position_ids)
    assert torch.equal(fused.cpu().view(Bn, S, nH, D).transpose(1, 2), want)
    sep, dummy = plain.clone(), plain.clone()
    lib = _cabi.lib()
    _cabi.check(lib.mc_rope(sep.data_ptr(), dummy.data_ptr(), cos_d.data_ptr(), sin_d.data_ptr(), Bn * S, S, 7, nH, D, H, H,
                            _cabi.dtype_code(dtype), _cabi.current_stream_ptr()), "rope")
    torch.cuda.synchronize()
    assert torch.equal(sep, fused)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("tuning", [0, 1, 3, 4])
def test_row_map_scatters_output_and_rope_positions(dtype, tuning):
    """modality-major row order: problem row m lands in row c_rowmap[m] of C, and RoPE takes the position of the mapped row;
    the result equals the un-permuted launch bit for bit (rows are independent)."""
    Bn, S, nH, D, K = 3, 50, 2, 128, 192
    T, H = Bn * S, nH * D
    g = torch.Generator().manual_seed(31)
    perm = torch.randperm(T, generator=g).to(torch.int32)
    x, W = rnd((T, K), dtype, 1).cuda(), rnd((H, K), dtype, 2, 0.1).cuda()
    cos, sin = (t.cuda() for t in XO.rope_cos_sin(D, 128, dtype))
    pos = torch.tensor([5], dtype=torch.int32, device="cuda")
    perm_d = perm.cuda()
    xp = torch.empty_like(x)
    LN.gather_rows(x, perm_d, xp)
    torch.cuda.synchronize()
    assert torch.equal(xp.cpu(), x.cpu()[perm.long()])
    for epi, rope in ((LN.EPI_NONE, None), (LN.EPI_ROPE, (cos, sin, pos, S, D))):
        want = torch.empty((T, H), dtype=dtype, device="cuda")
        LN.LinearPlan([LN.Problem(x, W, want, epilogue=epi, rope=rope)], tuning=tuning).run()
        got = torch.full((T, H), 7.0, dtype=dtype, device="cuda")
        LN.LinearPlan([LN.Problem(xp, W, got, epilogue=epi, rope=rope, c_rowmap=perm_d)], tuning=tuning).run()
        torch.cuda.synchronize()
        assert torch.equal(got, want), (epi, tuning)
    # strided gather (a column slice as source and destination)
    big = rnd((T, 512), dtype, 3).cuda()
    out = torch.zeros((T, 512), dtype=dtype, device="cuda")
    LN.gather_rows(big[:, 128:384], perm_d, out[:, 256:512])
    torch.cuda.synchronize()
    assert torch.equal(out[:, 256:512].cpu(), big.cpu()[perm.long(), 128:384]) and not out[:, :256].any()
    with pytest.raises(ValueError):
        LN.gather_rows(big, perm_d.long(), out)


def test_k_extension_unrouted():
    dtype = torch.bfloat16
    M, N, K0, K1 = 260, 512, 192, 128
    A0, B0 = rnd((M, K0), dtype, 1).cuda(), rnd((N, K0), dtype, 2, 0.1).cuda()
    A1, B1 = rnd((M, K1), dtype, 3).cuda(), rnd((N, K1), dtype, 4, 0.1).cuda()
    C = torch.empty((M, N), dtype=dtype, device="cuda")
    LN.LinearPlan([LN.Problem(A0, B0, C, A1=A1, B1=B1)]).run()
    check(C, A0.float() @ B0.float().t() + A1.float() @ B1.float().t(), dtype, "k-extension")


def make_modal_id(T, seed, n_groups, runs=True):
    g = torch.Generator().manual_seed(seed)
    if not runs:
        return torch.randint(0, n_groups, (T,), generator=g).to(torch.uint8)
    out, i = torch.zeros(T, dtype=torch.uint8), 0
    while i < T:
        n = int(torch.randint(1, 300, (1,), generator=g))
        out[i:i + n] = int(torch.randint(0, n_groups, (1,), generator=g))
        i += n
    return out


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("down_tuning", [0, 1, 1 | (1 << 16)])  # default tile / 128x128 compact schedule / 128x128 static
@pytest.mark.parametrize("T,in_f,out_f,r,runs", [(1000, 256, 512, 8, True), (700, 688, 256, 8, False), (1500, 512, 1024, 128, True),
                                                 (40000, 256, 256, 128, True)])
def test_routed_lora_linear_vs_oracle(dtype, T, in_f, out_f, r, runs, down_tuning):
    """LocalLoraLinear.forward + masked-sum routing (oracle, fp32 on the same 16-bit values) vs the two routed launches."""
    modal_names = ["default", "audio", "vision", "video"]
    dnames = [f"default-{m}" for m in modal_names[1:]]
    names = modal_names[1:] + dnames
    W = rnd((out_f, in_f), dtype, 1, 0.05)
    A = {n: rnd((r, in_f), dtype, 10 + i, 0.1) for i, n in enumerate(names)}
    Bm = {n: rnd((out_f, r), dtype, 30 + i, 0.1) for i, n in enumerate(names)}
    scaling = {n: 2.0 for n in names}
    for n in dnames:
        scaling[n] = 2.0 * 0.333
    scaling["default"] = 2.0
    x = rnd((T, in_f), dtype, 7)
    mid = make_modal_id(T, 5, len(modal_names), runs)
    # oracle in fp32
    f32 = lambda d: {k: v.float() for k, v in d.items()}
    outs = XO.lora_linear_forward(x.float()[None], W.float(), f32(A), f32(Bm), scaling, modal_names, dnames)
    masks = {m: (mid == i)[None] for i, m in enumerate(modal_names)}
    ref = XO.routed_sum(outs, masks, x.float()[None])[0]
    # CUDA path
    pk = LN.pack_adapters({k: v.cuda() for k, v in A.items()}, {k: v.cuda() for k, v in Bm.items()}, scaling, modal_names,
                          dnames, in_f, out_f, dtype, torch.device("cuda"))
    assert pk.group_cols[1] == LN.pad64(3 * r) and pk.A_all.shape[0] == pk.group_cols[-1]
    xd, rg = x.cuda(), mid.cuda()
    mt = LN.route_tile_masks(rg)
    Tb = torch.full((T, pk.A_all.shape[0]), float("nan"), dtype=dtype, device="cuda")  # skipped tiles must never be read
    y = torch.empty((T, out_f), dtype=dtype, device="cuda")
    LN.LinearPlan([LN.Problem(xd, pk.A_all, Tb, col_scale=pk.col_scale, row_group=rg, mtile_mask=mt, group_cols=pk.group_cols,
                              epilogue=LN.EPI_ROWMASK)], tuning=down_tuning).run()
    LN.LinearPlan([LN.Problem(xd, W.cuda(), y, A1=Tb, B1=pk.B_all, mtile_mask=mt, group_cols=pk.group_cols)]).run()
    torch.cuda.synchronize()
    assert not torch.isnan(y).any()
    # T is rounded to 16 bits between the two launches (the reference rounds there too): allow one extra ulp of scale
    got, ref = y.float().cpu(), ref
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    assert (err <= RTOL[dtype] * ref.abs() + RTOL[dtype] * scale).all(), err.max().item() / scale


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("up_tuning", [3, 4])
@pytest.mark.parametrize("runs", [True, False])
def test_routed_up_launch_on_pair_kernels_is_bit_identical(dtype, up_tuning, runs):
    """The CTA-pair kernels skip LoRA k-blocks on the union of the groups of 256 / 512 rows: the extra blocks only multiply
    zeros and the k-blocks are accumulated in the same order, so y equals the single-CTA kernel's bit for bit.  Three
    problems per launch, residual epilogue, ragged M."""
    T, in_f, out_f, r = 3000, 512, 768, 64
    modal_names = ["default", "audio", "vision"]
    dnames = [f"default-{m}" for m in modal_names[1:]]
    names = modal_names[1:] + dnames
    A = {n: rnd((r, in_f), dtype, 10 + i, 0.1).cuda() for i, n in enumerate(names)}
    Bm = {n: rnd((out_f, r), dtype, 30 + i, 0.1).cuda() for i, n in enumerate(names)}
    scaling = {n: 2.0 for n in names + ["default"]}
    pk = LN.pack_adapters(A, Bm, scaling, modal_names, dnames, in_f, out_f, dtype, torch.device("cuda"))
    x, res = rnd((T, in_f), dtype, 7).cuda(), rnd((T, out_f), dtype, 8).cuda()
    Ws = [rnd((out_f, in_f), dtype, 40 + i, 0.05).cuda() for i in range(3)]
    rg = make_modal_id(T, 5, len(modal_names), runs).cuda()
    outs = {}
    for tuning in (0, up_tuning):
        # the pair kernels read the rank columns of every group present in 256 / 512 rows: the tile masks are coarsened to
        # that granularity so the down-projection writes them (T starts as NaN: a column it skipped must never be read)
        mt = LN.route_tile_masks(rg, coarsen={0: 1, 3: 4, 4: 2}[tuning])
        Tb = torch.full((T, pk.A_all.shape[0]), float("nan"), dtype=dtype, device="cuda")
        LN.LinearPlan([LN.Problem(x, pk.A_all, Tb, col_scale=pk.col_scale, row_group=rg, mtile_mask=mt, group_cols=pk.group_cols,
                                  epilogue=LN.EPI_ROWMASK)], tuning=1).run()
        ys = [torch.empty((T, out_f), dtype=dtype, device="cuda") for _ in range(3)]
        LN.LinearPlan([LN.Problem(x, W, y, A1=Tb, B1=pk.B_all, mtile_mask=mt, group_cols=pk.group_cols, residual=res,
                                  epilogue=LN.EPI_RESIDUAL) for W, y in zip(Ws, ys)], tuning=tuning).run()
        torch.cuda.synchronize()
        assert not any(torch.isnan(y).any() for y in ys)
        outs[tuning] = ys
    for a, b in zip(outs[0], outs[up_tuning]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("T,n_groups,lut", [(1, 1, None), (1000, 4, None), (31360, 4, [0, 2, 1, 3, 0, 0, 0]), (57104, 5, [0, 4, 3, 2, 1, 0, 0]), (777, 8, None)])
def test_route_permutation_is_a_stable_sort_by_group(T, n_groups, lut):
    g = torch.Generator().manual_seed(T)
    n_ids = n_groups if lut is None else 5
    ids = torch.randint(0, n_ids, (T,), generator=g).to(torch.uint8).cuda()
    perm, inv = torch.empty(T, dtype=torch.int32, device="cuda"), torch.empty(T, dtype=torch.int32, device="cuda")
    rg, seq = torch.empty(T, dtype=torch.uint8, device="cuda"), torch.empty(T, dtype=torch.uint8, device="cuda")
    seg = torch.empty(n_groups + 1, dtype=torch.int32, device="cuda")
    LN.route_permutation(ids, lut, n_groups, perm, inv, rg, seg, seq)
    groups = ids.long().cpu() if lut is None else torch.tensor(lut)[ids.long().cpu()]
    order = torch.sort(groups, stable=True).indices
    assert torch.equal(perm.cpu().long(), order) and torch.equal(seq.cpu().long(), groups)
    assert torch.equal(inv.cpu().long()[order], torch.arange(T)) and torch.equal(rg.cpu().long(), groups[order])
    counts = torch.bincount(groups, minlength=n_groups)
    assert torch.equal(seg.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))


def test_route_tile_masks():
    mid = make_modal_id(1000, 3, 5)
    got = LN.route_tile_masks(mid.cuda()).cpu()
    for t in range(got.numel()):
        want = 0
        for g in mid[t * 128:(t + 1) * 128].unique().tolist():
            want |= 1 << g
        assert got[t].item() == want
    for coarsen in (2, 4):  # every tile of an aligned run carries the union of the run
        c = LN.route_tile_masks(mid.cuda(), coarsen=coarsen).cpu()
        for t in range(got.numel()):
            want = 0
            for u in range(t // coarsen * coarsen, min(got.numel(), t // coarsen * coarsen + coarsen)):
                want |= got[u].item()
            assert c[t].item() == want, (coarsen, t)


@pytest.mark.parametrize("key", ["torch.bfloat16", "torch.float16"])
def test_projector_grouped_vs_reference_fixture(golden, key):
    """Both modalities' mlp2x_gelu projectors as ONE grouped launch per layer vs the reference's own outputs."""
    dtype = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}[key]
    m = golden("merge_c1.pt")["runs"]["online-merge-reset-default-vision=0.5,default-audio=0.5"]["state_dict"]
    g = golden("projector.pt")
    probs1, probs2, outs = [], [], {}
    for modal in ("vision", "audio"):
        pre = f"model.modal_projectors.{modal}."
        x = g[modal]["x"].to(dtype).cuda()
        x2 = x.view(-1, x.shape[-1])
        W0, b0, W2, b2 = (m[pre + k].to(dtype).cuda() for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
        h = torch.empty((x2.shape[0], W0.shape[0]), dtype=dtype, device="cuda")
        o = torch.empty((x2.shape[0], W2.shape[0]), dtype=dtype, device="cuda")
        probs1.append(LN.Problem(x2, W0, h, bias=b0, epilogue=LN.EPI_BIAS_GELU))
        probs2.append(LN.Problem(h, W2, o, bias=b2, epilogue=LN.EPI_BIAS))
        outs[modal] = (o, x.shape)
    LN.LinearPlan(probs1).run()
    LN.LinearPlan(probs2).run()
    torch.cuda.synchronize()
    tol = {"torch.bfloat16": 2 ** -6, "torch.float16": 2 ** -9}[key]
    for modal, (o, shp) in outs.items():
        ref = g[modal][key].float()
        got = o.view(shp[0], shp[1], -1).float().cpu()
        scale = ref.abs().max().item()
        assert (got - ref).abs().max().item() <= tol * scale, modal
        # and against the oracle evaluated in fp32
        pre = f"model.modal_projectors.{modal}."
        want = XO.projector_forward(g[modal]["x"].to(dtype).float(), [m[pre + "0.weight"].to(dtype).float(), m[pre + "2.weight"].to(dtype).float()],
                                    [m[pre + "0.bias"].to(dtype).float(), m[pre + "2.bias"].to(dtype).float()])
        assert (got - want).abs().max().item() <= tol * scale


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_rowwise_glue_bit_exact_vs_oracle(dtype):
    """RMSNorm / RoPE / SiLU·mul restate the eager reference ops with the same rounding points."""
    T, H, nH = 77, 256, 4
    D = H // nH
    x = rnd((T, H), dtype, 1)
    w = (1 + rnd((H,), torch.float32, 2, 0.02)).to(dtype)
    lib = _cabi.lib()
    code, st = _cabi.dtype_code(dtype), _cabi.current_stream_ptr()
    xd, wd = x.cuda(), w.cuda()
    out = torch.empty_like(xd)
    _cabi.check(lib.mc_rmsnorm(xd.data_ptr(), wd.data_ptr(), out.data_ptr(), T, H, H, H, 1e-5, code, st), "rmsnorm")
    ref = XO.rms_norm(x, w, 1e-5)
    # fp32 sum order differs (warp tree vs torch): allow one ulp on a handful of elements, none elsewhere
    diff = (out.cpu().float() - ref.float()).abs()
    assert (diff <= RTOL[dtype] * 2 * ref.float().abs() + 1e-6).all()
    assert (diff > 0).float().mean().item() < 0.02
    # rope: S = 11 positions, 7 sequences
    S, Bn = 11, 7
    q, k = rnd((Bn * S, H), dtype, 3), rnd((Bn * S, H), dtype, 4)
    cos, sin = XO.rope_cos_sin(D, 64, dtype)
    qd, kd, cos_d, sin_d = q.cuda(), k.cuda(), cos.cuda(), sin.cuda()
    _cabi.check(lib.mc_rope(qd.data_ptr(), kd.data_ptr(), cos_d.data_ptr(), sin_d.data_ptr(), Bn * S, S, 0, nH, D, H, H,
                            code, st), "rope")
    pos = torch.arange(S)[None].expand(Bn, S)
    qr, kr = XO.apply_rope(q.view(Bn, S, nH, D).transpose(1, 2), k.view(Bn, S, nH, D).transpose(1, 2), cos, sin, pos)
    assert torch.equal(qd.cpu().view(Bn, S, nH, D).transpose(1, 2), qr)
    assert torch.equal(kd.cpu().view(Bn, S, nH, D).transpose(1, 2), kr)
    # decode-step form: one token per sequence at position 9
    q1, k1 = q.view(Bn, S, H)[:, 9].contiguous(), k.view(Bn, S, H)[:, 9].contiguous()
    q1d, k1d = q1.cuda(), k1.cuda()
    _cabi.check(lib.mc_rope(q1d.data_ptr(), k1d.data_ptr(), cos_d.data_ptr(), sin_d.data_ptr(), Bn, 1, 9, nH, D, H, H, code, st), "rope")
    assert torch.equal(q1d.cpu().view(Bn, nH, D), qr[:, :, 9]) and torch.equal(k1d.cpu().view(Bn, nH, D), kr[:, :, 9])
    # silu * mul
    g_, u_ = rnd((T, 688), dtype, 5, 2.0), rnd((T, 688), dtype, 6)
    got = LN.silu_mul(g_.cuda(), u_.cuda()).cpu()
    want = torch.nn.functional.silu(g_) * u_
    d = (got.float() - want.float()).abs()
    assert (d <= RTOL[dtype] * 2 * want.float().abs() + 1e-7).all() and (d > 0).float().mean().item() < 0.01


def test_argument_errors():
    a = torch.zeros((8, 8), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        LN.LinearPlan([LN.Problem(a.cpu(), a, a)])
    with pytest.raises(_cabi.McError, match="multiples of 8"):
        LN.LinearPlan([LN.Problem(torch.zeros((8, 16), dtype=torch.bfloat16, device="cuda")[:, :12],
                                  torch.zeros((8, 12), dtype=torch.bfloat16, device="cuda"), a)])
    with pytest.raises(_cabi.McError, match="ROWMASK"):
        LN.LinearPlan([LN.Problem(a, a, a.clone(), epilogue=LN.EPI_ROWMASK)])
    with pytest.raises(_cabi.McError, match="dtype"):
        LN.LinearPlan([LN.Problem(a.float(), a.float(), a.float())])
    # ROPE epilogue: 32-byte vectors — an output view that is only 16-byte aligned is refused at plan creation
    x, W = torch.zeros((16, 64), dtype=torch.bfloat16, device="cuda"), torch.zeros((128, 64), dtype=torch.bfloat16, device="cuda")
    cos, sin = (t.cuda() for t in XO.rope_cos_sin(64, 32, torch.bfloat16))
    pos = torch.zeros(1, dtype=torch.int32, device="cuda")
    wide = torch.zeros((16, 144), dtype=torch.bfloat16, device="cuda")
    LN.LinearPlan([LN.Problem(x, W, wide[:, :128], epilogue=LN.EPI_ROPE, rope=(cos, sin, pos, 16, 64))])   # aligned view: accepted
    with pytest.raises(_cabi.McError, match="32-byte"):
        LN.LinearPlan([LN.Problem(x, W, wide[:, 8:136], epilogue=LN.EPI_ROPE, rope=(cos, sin, pos, 16, 64))])
    with pytest.raises(_cabi.McError, match="32-byte"):
        LN.LinearPlan([LN.Problem(x, W, torch.zeros((16, 136), dtype=torch.bfloat16, device="cuda")[:, :128], epilogue=LN.EPI_ROPE,
                                  rope=(cos, sin, pos, 16, 64))])   # ldc = 136: not a multiple of 16
