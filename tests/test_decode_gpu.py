"""GPU parity of the decode-step kernels (csrc/mc_decode.cu) and of the graph-captured decode loop.

Kernels are compared with an fp32 PyTorch evaluation of the same op on the same 16-bit inputs (tolerance: one rounding of the
fp32 result to the storage dtype, 2^-8 relative for bf16 / 2^-11 for fp16, plus fp32 summation-order noise); the model-level
tests compare decode steps with a from-scratch prefill of the extended sequence (multimodal_arch.py:290-293,
multimodal_llama.py:436-438 semantics: cache present -> default adapter for every row), which tests/test_prefill_gpu.py pins
to the reference fixtures and the oracle."""
import math

import pytest
import torch

from modelcompose_b200 import _cabi
from modelcompose_b200 import decode as DC
from modelcompose_b200 import model as MD
from modelcompose_b200 import synthetic as syn
from oracle import splice_oracle as SO

pytestmark = pytest.mark.gpu

DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16}
ULP = {"bf16": 2.0 ** -8, "fp16": 2.0 ** -11}
STRATEGY_C1 = "online-merge-reset-default-vision=0.5,default-audio=0.5"


def close(got, ref, key, what, extra=0.0):
    got, ref = got.float().cpu(), ref.float().cpu()
    tol = ULP[key] * ref.abs() + (ULP[key] / 2 + extra) * ref.abs().max()
    bad = (got - ref).abs() > tol
    assert not bad.any(), (what, (got - ref).abs().max().item(), ref.abs().max().item(), int(bad.sum()))


def rnd(shape, g, dtype, std=1.0):
    return (torch.randn(shape, generator=g) * std).to(dtype).cuda()


# ------------------------------------------------------------------------------------------------------ skinny linear
@pytest.mark.parametrize("key", list(DTYPES))
@pytest.mark.parametrize("tuning", [0, 16, 17, 32, 128, 3 << 8])
@pytest.mark.parametrize("M,N,K0,K1", [(1, 64, 128, 0), (5, 200, 256, 64), (8, 4096, 1024, 384), (16, 136, 688, 48), (17, 256, 4096, 0),
                                       (32, 512, 2048, 384), (33, 328, 640, 128), (64, 1024, 1024, 64), (3, 72, 72, 24)])
def test_skinny_linear_vs_fp32(key, tuning, M, N, K0, K1):
    dt = DTYPES[key]
    g = torch.Generator().manual_seed(M * 1000 + N + K0 + K1)
    x, W = rnd((M, K0), g, dt), rnd((N, K0), g, dt, K0 ** -0.5)
    ref = x.float() @ W.float().t()
    p = dict(A0=x, B0=W)
    if K1:
        t, Bu = rnd((M, K1), g, dt), rnd((N, 2 * K1), g, dt, K1 ** -0.5)[:, :K1]  # strided B1 as B_all[:, :R0]
        p.update(A1=t, B1=Bu)
        ref = ref + t.float() @ Bu.float().t()
    out = torch.full((M, N), float("nan"), dtype=dt, device="cuda")
    DC.SkinnyLaunch([dict(p, C=out)], tuning).run()
    close(out, ref, key, "plain")
    res = rnd((M, N), g, dt)
    out2 = res.clone()
    DC.SkinnyLaunch([dict(p, C=out2, residual=out2, epilogue=DC.SK_RESIDUAL)], tuning).run()
    close(out2, ref + res.float(), key, "residual in place")
    cs = torch.rand(N, generator=g).cuda() * 2
    out3 = torch.empty_like(out)
    DC.SkinnyLaunch([dict(p, C=out3, col_scale=cs, epilogue=DC.SK_COLSCALE)], tuning).run()
    close(out3, ref * cs.float().cpu().cuda()[None], key, "colscale")
    torch.cuda.synchronize()
    # the workspace must be left zeroed (graph replays depend on it)
    assert int(DC.skinny_workspace("cuda")[:4 * 8192].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("key", list(DTYPES))
@pytest.mark.parametrize("tuning", [0, 16, 32, 128])
@pytest.mark.parametrize("M,N,K0,K1", [(4, 96, 256, 0), (32, 11008, 4096, 384), (20, 688, 256, 16), (64, 344, 512, 64)])
def test_skinny_dual_silu_mul(key, tuning, M, N, K0, K1):
    """gate/up in one launch: silu(gate) * up with gate, silu(gate), up each rounded to the storage dtype (the prefill's rounding points)."""
    dt = DTYPES[key]
    g = torch.Generator().manual_seed(N + K0)
    x, Wg, Wu = rnd((M, K0), g, dt), rnd((N, K0), g, dt, K0 ** -0.5), rnd((N, K0), g, dt, K0 ** -0.5)
    gate, up = x.float() @ Wg.float().t(), x.float() @ Wu.float().t()
    p = dict(A0=x, B0=Wg, B0u=Wu, epilogue=DC.SK_SILU_MUL)
    if K1:
        tg, tu = rnd((M, K1), g, dt), rnd((M, K1), g, dt)
        Bg, Bu = rnd((N, K1), g, dt, K1 ** -0.5), rnd((N, K1), g, dt, K1 ** -0.5)
        p.update(A1=tg, B1=Bg, A1u=tu, B1u=Bu)
        gate, up = gate + tg.float() @ Bg.float().t(), up + tu.float() @ Bu.float().t()
    out = torch.empty((M, N), dtype=dt, device="cuda")
    DC.SkinnyLaunch([dict(p, C=out)], tuning).run()
    gate_r = gate.to(dt).float()
    ref = torch.nn.functional.silu(gate_r).to(dt).float() * up.to(dt).float()
    # one-ulp differences of the rounded gate move silu by up to an ulp of the product: 3 ulp budget
    close(out, ref, key, "silu*mul", extra=2 * ULP[key])


def test_skinny_multi_problem_qkv_and_kernel_agreement():
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    M, H, R0 = 32, 1024, 128
    x = rnd((M, H), g, dt)
    Ws = [rnd((H, H), g, dt, H ** -0.5) for _ in range(3)]
    ts = [rnd((M, R0), g, dt) for _ in range(3)]
    Bs = [rnd((H, R0), g, dt, R0 ** -0.5) for _ in range(3)]
    outs = {}
    for tuning in (0, 16, 32, 128):
        o = [torch.empty((M, H), dtype=dt, device="cuda") for _ in range(3)]
        DC.SkinnyLaunch([dict(A0=x, B0=Ws[i], A1=ts[i], B1=Bs[i], C=o[i]) for i in range(3)], tuning).run()
        outs[tuning] = o
        for i in range(3):
            close(o[i], x.float() @ Ws[i].float().t() + ts[i].float() @ Bs[i].float().t(), "bf16", f"problem {i} tuning {tuning}")
    # deterministic: a second run of the same launch is bit-identical (fixed-order combine of split row blocks)
    o2 = [torch.empty((M, H), dtype=dt, device="cuda") for _ in range(3)]
    DC.SkinnyLaunch([dict(A0=x, B0=Ws[i], A1=ts[i], B1=Bs[i], C=o2[i]) for i in range(3)], 0).run()
    assert all(torch.equal(a, b) for a, b in zip(outs[0], o2))


def test_skinny_rejects_bad_arguments():
    x = torch.zeros((65, 64), dtype=torch.bfloat16, device="cuda")
    W = torch.zeros((64, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        DC.SkinnyLaunch([dict(A0=x, B0=W, C=torch.zeros((65, 64), dtype=torch.bfloat16, device="cuda"))])
    with pytest.raises(ValueError):
        DC.SkinnyLaunch([dict(A0=x[:4].cpu(), B0=W, C=x[:4])])


# ------------------------------------------------------------------------------------------------------ attention
def _rope_tables(n, D, dt):
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
    f = torch.einsum("i,j->ij", torch.arange(n).float(), inv)
    emb = torch.cat((f, f), -1)
    return emb.cos().to(dt).cuda(), emb.sin().to(dt).cuda()


@pytest.mark.parametrize("key", list(DTYPES))
@pytest.mark.parametrize("B,nH,L,splits,masked", [(1, 2, 1, 1, False), (3, 4, 37, 1, False), (2, 32, 300, 4, False), (4, 8, 1000, 16, True),
                                                  (32, 4, 129, 2, True), (2, 2, 64, 3, False)])
def test_decode_rope_append_and_attention(key, B, nH, L, splits, masked):
    dt, D = DTYPES[key], 128
    g = torch.Generator().manual_seed(B + L)
    cap = L + 7
    kc, vc = rnd((B, nH, cap, D), g, dt), rnd((B, nH, cap, D), g, dt)   # [batch, heads, capacity, D]
    q, k, v = rnd((B, nH * D), g, dt), rnd((B, nH * D), g, dt), rnd((B, nH * D), g, dt)
    q_raw = q.clone()
    cos, sin = _rope_tables(cap + 8, D, dt)
    pos = torch.tensor([L - 1], dtype=torch.int32, device="cuda")
    # reference RoPE through the library's own prefill op (bit-exact expected: same rounding points)
    q_ref, k_ref = q.clone(), k.clone()
    _cabi.check(_cabi.lib().mc_rope(q_ref.data_ptr(), k_ref.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, 1, L - 1, nH, D, nH * D, nH * D,
                                    _cabi.dtype_code(dt), _cabi.current_stream_ptr()), "mc_rope")
    lib, st = _cabi.lib(), _cabi.current_stream_ptr()
    _cabi.check(lib.mc_decode_rope_append(q.data_ptr(), k.data_ptr(), v.data_ptr(), nH * D, kc.data_ptr(), vc.data_ptr(), cap, pos.data_ptr(),
                                          cos.data_ptr(), sin.data_ptr(), B, nH, D, _cabi.dtype_code(dt), st), "rope_append")
    assert torch.equal(q, q_ref)
    assert torch.equal(kc[:, :, L - 1].reshape(B, -1), k_ref) and torch.equal(vc[:, :, L - 1].reshape(B, -1), v)
    mask = None
    if masked:
        mask = (torch.rand((B, cap), generator=g) > 0.3).to(torch.uint8)
        mask[:, L - 1] = 1
        mask = mask.cuda()
    out = torch.empty((B, nH * D), dtype=dt, device="cuda")
    scratch = torch.zeros(B * nH * splits * (D + 2), dtype=torch.float32, device="cuda")
    counters = torch.zeros(B * nH, dtype=torch.int32, device="cuda")
    for _ in range(2):  # twice: the counters must come back to zero
        _cabi.check(lib.mc_decode_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), cap, pos.data_ptr(),
                                            None if mask is None else mask.data_ptr(), 0 if mask is None else cap, out.data_ptr(),
                                            nH * D, nH * D, B, nH, D, 1.0 / math.sqrt(D), splits, scratch.data_ptr(), counters.data_ptr(),
                                            _cabi.dtype_code(dt), st), "decode_attention")
    assert int(counters.abs().sum()) == 0
    # the fused launch (RoPE + append + attention) from the raw q / k / v: same cache contents and output, bit for bit
    kc2, vc2 = kc.clone(), vc.clone()
    kc2[:, :, L - 1], vc2[:, :, L - 1] = 7.0, -7.0
    out2 = torch.empty_like(out)
    q_raw_in = q_raw.clone()
    _cabi.check(lib.mc_decode_attention_fused(q_raw.data_ptr(), k.data_ptr(), v.data_ptr(), nH * D, kc2.data_ptr(), vc2.data_ptr(), cap, pos.data_ptr(),
                                              cos.data_ptr(), sin.data_ptr(), None if mask is None else mask.data_ptr(), 0 if mask is None else cap,
                                              out2.data_ptr(), nH * D, B, nH, D, 1.0 / math.sqrt(D), splits, scratch.data_ptr(), counters.data_ptr(),
                                              _cabi.dtype_code(dt), st), "decode_attention_fused")
    assert torch.equal(kc2, kc) and torch.equal(vc2, vc) and torch.equal(out2, out) and torch.equal(q_raw, q_raw_in)
    assert int(counters.abs().sum()) == 0
    qf = q.float().view(B, nH, 1, D)
    kf, vf = kc[:, :, :L].float(), vc[:, :, :L].float()
    s = (qf @ kf.transpose(-1, -2)) / math.sqrt(D)
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :L] == 0, float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, nH * D)
    close(out, ref, key, "decode attention", extra=ULP[key])


@pytest.mark.parametrize("key", list(DTYPES))
def test_argmax_rows(key):
    dt = DTYPES[key]
    g = torch.Generator().manual_seed(3)
    lg = rnd((7, 32000), g, dt)
    lg[2, 100] = lg[2, 31999] = 50.0   # tie: first index wins
    lg[3, 31999] = 60.0
    lg[4, 0] = 60.0
    o32 = torch.empty(7, dtype=torch.int32, device="cuda")
    o64 = torch.empty(7, dtype=torch.int64, device="cuda")
    counter = torch.tensor([41], dtype=torch.int32, device="cuda")
    DC.argmax_rows(lg, o32, o64, counter)
    want = torch.stack([(row == row.max()).nonzero()[0, 0] for row in lg.float()])
    assert torch.equal(o64, want) and torch.equal(o32.long(), want) and int(counter) == 42
    assert int(o64[2]) == 100 and int(o64[3]) == 31999 and int(o64[4]) == 0


# ------------------------------------------------------------------------------------------------------ model level
def d128_model(golden, dtype, materialize=False):
    """The C1 composed tiny model viewed with 2 heads of 128 (same weights): the decode / attention kernels need head_dim 128."""
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfgd = dict(run["config"])
    cfgd["num_attention_heads"] = cfgd["num_key_value_heads"] = 2
    base = syn.make_base_llm(seed=1)
    return MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfgd), base, run["state_dict"], device="cuda", dtype=dtype,
                                         materialize=materialize)


def _prompt(B, dtype, seed=8):
    g = torch.Generator().manual_seed(seed + 13)
    ids = syn.make_prompt_ids(B, ["vision", "audio"], 9, 1000, seed=seed, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=4)
    feats = {"audio": torch.randn(B, 6, 48, generator=g).to(dtype).cuda(), "vision": torch.randn(B, 11, 64, generator=g).to(dtype).cuda()}
    return ids.cuda(), feats


MAXABS = {"bf16": 2 ** -5, "fp16": 2 ** -8}
COS = {"bf16": 0.9995, "fp16": 0.99999}


def logits_close(got, ref, key, what):
    got, ref = got.float().cpu().flatten(), ref.float().cpu().flatten()
    scale, err = ref.abs().max().item(), (got - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    print(f"{what} [{key}]: max-abs {err:.4g} (scale {scale:.4g}) cosine {cos:.7f}")
    assert err <= MAXABS[key] * scale and cos >= COS[key], (what, err, scale, cos)


@pytest.mark.parametrize("key", list(DTYPES))
@pytest.mark.parametrize("materialize", [False, True])
def test_decode_steps_vs_prefill_of_extended_sequence(golden, key, materialize):
    dtype = DTYPES[key]
    model = d128_model(golden, dtype, materialize)
    B, new = 3, 5
    ids, feats = _prompt(B, dtype)
    out_ids = model.generate(ids, modal_inputs=feats, max_new_tokens=new, do_sample=False)
    assert out_ids.shape == (B, ids.shape[1] + new) and torch.equal(out_ids[:, :ids.shape[1]], ids)
    assert model._dws is not None and model._dws.graph is not None  # the captured step ran
    full = model.forward(out_ids, torch.ones_like(out_ids), modal_inputs=feats)
    Sp = full.logits.shape[1]
    # every generated token is (within tolerance) the argmax of the teacher-forced prefill at its position
    for i in range(new):
        step_logits = full.logits[:, Sp - new + i - 1, :].float()
        chosen = step_logits.gather(1, out_ids[:, ids.shape[1] + i][:, None]).squeeze(1)
        assert ((step_logits.max(-1).values - chosen) <= MAXABS[key] * step_logits.abs().max()).all(), i
    # explicit decode steps through forward(past_key_values=...): logits of steps 1..3 vs the prefill at the same positions
    o1 = model.forward(ids, torch.ones_like(ids), modal_inputs=feats, use_cache=True, cache_extra=8)
    cache = o1.past_key_values
    S0 = cache.length
    for i in range(3):
        tok = out_ids[:, ids.shape[1] + i:ids.shape[1] + i + 1]
        o = model.forward(tok, torch.ones((B, cache.length + 1), dtype=torch.int64, device="cuda"), past_key_values=cache, modal_inputs=feats)
        assert cache.length == S0 + i + 1 and o.logits.shape == (B, 1, 1000)
        logits_close(o.logits[:, 0], full.logits[:, Sp - new + i, :], key, f"decode step {i} vs prefill")


def test_decode_with_linear_rope_scaling(golden):
    """config.rope_scaling "linear": the decode step reads the same scaled table as the prefill (decode vs prefill of the extended sequence)."""
    dtype, key = torch.bfloat16, "bf16"
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfgd = dict(run["config"])
    cfgd["num_attention_heads"] = cfgd["num_key_value_heads"] = 2
    cfgd["rope_scaling"] = {"type": "linear", "factor": 4.0}
    model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfgd), syn.make_base_llm(seed=1), run["state_dict"], device="cuda", dtype=dtype)
    B, new = 2, 4
    ids, feats = _prompt(B, dtype)
    out_ids = model.generate(ids, modal_inputs=feats, max_new_tokens=new, do_sample=False)
    full = model.forward(out_ids, torch.ones_like(out_ids), modal_inputs=feats)
    Sp = full.logits.shape[1]
    o1 = model.forward(ids, torch.ones_like(ids), modal_inputs=feats, use_cache=True, cache_extra=8)
    cache = o1.past_key_values
    for i in range(3):
        tok = out_ids[:, ids.shape[1] + i:ids.shape[1] + i + 1]
        o = model.forward(tok, torch.ones((B, cache.length + 1), dtype=torch.int64, device="cuda"), past_key_values=cache, modal_inputs=feats)
        logits_close(o.logits[:, 0], full.logits[:, Sp - new + i, :], key, f"decode step {i} vs prefill, linear RoPE scaling")


def test_decode_dense_option(golden):
    """decode_dense: branch form for the prefill, the dense W_eff of the text group for the decode steps only — the same weights the
    materialised form holds for group 0 (bit for bit), four launches per layer fewer, decode logits within the bar of the prefill."""
    dtype, key = torch.bfloat16, "bf16"
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfgd = dict(run["config"])
    cfgd["num_attention_heads"] = cfgd["num_key_value_heads"] = 2
    cfg, base = MD.MultimodalConfig.from_dict(cfgd), syn.make_base_llm(seed=1)
    model = MD.MultimodalLlamaForCausalLM(cfg, base, run["state_dict"], device="cuda", dtype=dtype, decode_dense=True)
    dense = d128_model(golden, dtype, materialize=True)
    branch = d128_model(golden, dtype)
    assert model.decode_dense and not model.materialize and not dense.decode_dense
    for la, lb in zip(model.layers, dense.layers):
        for n in la.Wdec:
            assert torch.equal(la.Wdec[n], lb.Weff[n][0]), n
    B, new = 3, 4
    ids, feats = _prompt(B, dtype)
    out_ids = model.generate(ids, modal_inputs=feats, max_new_tokens=new, do_sample=False)
    branch.generate(ids, modal_inputs=feats, max_new_tokens=new, do_sample=False)
    assert model._dws.launches_per_step() == branch._dws.launches_per_step() - 4 * len(model.layers)
    full = model.forward(out_ids, torch.ones_like(out_ids), modal_inputs=feats)   # branch-form prefill of the extended sequence
    Sp = full.logits.shape[1]
    o1 = model.forward(ids, torch.ones_like(ids), modal_inputs=feats, use_cache=True, cache_extra=8)
    cache = o1.past_key_values
    for i in range(3):
        tok = out_ids[:, ids.shape[1] + i:ids.shape[1] + i + 1]
        o = model.forward(tok, torch.ones((B, cache.length + 1), dtype=torch.int64, device="cuda"), past_key_values=cache, modal_inputs=feats)
        logits_close(o.logits[:, 0], full.logits[:, Sp - new + i, :], key, f"dense decode step {i} vs branch-form prefill")


def test_decode_graph_equals_eager_and_prefill_kernel_path(golden, monkeypatch):
    dtype = torch.bfloat16
    ids, feats = _prompt(4, dtype, seed=5)
    outs = {}
    for name, env in (("graph", {}), ("eager", {"DECODE_GRAPH": False}), ("prefill-kernels", {"DECODE_NATIVE": False})):
        for k, v in env.items():
            monkeypatch.setattr(MD, k, v)
        model = d128_model(golden, dtype)
        o1 = model.forward(ids, torch.ones_like(ids), modal_inputs=feats, use_cache=True, cache_extra=8)
        cache, logits = o1.past_key_values, []
        tok = o1.logits[:, -1].argmax(-1)
        for _ in range(4):
            o = model.forward(tok[:, None], None, past_key_values=cache, modal_inputs=feats)
            logits.append(o.logits[:, 0].clone())
            tok = o.logits[:, 0].argmax(-1)
        outs[name] = torch.stack(logits)
        monkeypatch.undo()
    assert torch.equal(outs["graph"], outs["eager"])
    logits_close(outs["graph"], outs["prefill-kernels"], "bf16", "decode kernels vs the prefill kernels at M = batch")


def test_programmatic_dependent_launch_is_bit_identical(golden, monkeypatch):
    """The decode chain launched with programmatic dependent launch (kernels overlap their predecessors' tails, weights
    prefetched before griddepcontrol.wait) against plain stream-ordered launches: same bits, graph and eager."""
    dtype = torch.bfloat16
    ids, feats = _prompt(4, dtype, seed=6)
    outs = {}
    for pdl in (True, False):
        for graph in (True, False):
            monkeypatch.setattr(DC, "PDL", pdl)
            monkeypatch.setattr(MD, "DECODE_GRAPH", graph)
            model = d128_model(golden, dtype)
            outs[(pdl, graph)] = model.generate(ids, modal_inputs=feats, max_new_tokens=12, do_sample=False)
            logits = model._dws.logits.clone()
            outs[(pdl, graph, "logits")] = logits
            assert model._dws.pdl == pdl and (model._dws.graph is not None) == graph
            monkeypatch.undo()
    for graph in (True, False):
        assert torch.equal(outs[(True, graph)], outs[(False, graph)])
        assert torch.equal(outs[(True, graph, "logits")], outs[(False, graph, "logits")])


def test_cache_growth_rebuilds_the_decode_workspace(golden):
    """forward(past_key_values=...) past the cache's capacity: the cache is reallocated, the captured graph (which bakes the cache
    addresses) is rebuilt, and the logits equal those of a run whose cache was large enough from the start."""
    dtype = torch.bfloat16
    model = d128_model(golden, dtype)
    ids, feats = _prompt(2, dtype, seed=4)
    runs = {}
    for extra in (2, 64):
        o1 = model.forward(ids, torch.ones_like(ids), modal_inputs=feats, use_cache=True, cache_extra=extra)
        cache, tok, logits = o1.past_key_values, o1.logits[:, -1].argmax(-1), []
        cap0 = cache.capacity
        for _ in range(6):
            o = model.forward(tok[:, None], None, past_key_values=cache, modal_inputs=feats)
            logits.append(o.logits[:, 0].clone())
            tok = o.logits[:, 0].argmax(-1)
        runs[extra] = torch.stack(logits)
        assert (cache.capacity > cap0) == (extra == 2)
    assert torch.equal(runs[2], runs[64])


def test_sampling_and_eos_on_the_native_decode_path(golden):
    dtype = torch.bfloat16
    model = d128_model(golden, dtype)
    ids, feats = _prompt(3, dtype, seed=3)
    greedy = model.generate(ids, modal_inputs=feats, max_new_tokens=5, do_sample=False)
    eos = int(greedy[0, ids.shape[1] + 1])  # row 0 emits it at its second step
    out = model.generate(ids, modal_inputs=feats, max_new_tokens=5, do_sample=False, eos_token_id=eos, pad_token_id=0)
    assert torch.equal(out[0, :ids.shape[1] + 2], greedy[0, :ids.shape[1] + 2]) and (out[0, ids.shape[1] + 2:] == 0).all()
    a = model.generate(ids, modal_inputs=feats, max_new_tokens=4, do_sample=True, temperature=0.8, top_p=0.9,
                       generator=torch.Generator(device="cuda").manual_seed(5))
    b = model.generate(ids, modal_inputs=feats, max_new_tokens=4, do_sample=True, temperature=0.8, top_p=0.9,
                       generator=torch.Generator(device="cuda").manual_seed(5))
    assert a.shape == (3, ids.shape[1] + 4) and torch.equal(a, b)


@pytest.mark.parametrize("what", ["FUSED_ROPE", "FUSED_NORM"])
@pytest.mark.parametrize("materialize", [False, True])
def test_fused_rope_attention_launch_is_bit_identical_in_the_loop(golden, monkeypatch, what, materialize):
    """RoPE + append inside the attention launch, and the RMSNorms inside the skinny launches that consume them, against the
    separate launches: same tokens, logits and cache contents, bit for bit."""
    dtype = torch.bfloat16
    ids, feats = _prompt(4, dtype, seed=12)
    outs = {}
    for fused in (True, False):
        monkeypatch.setattr(DC, what, fused)
        model = d128_model(golden, dtype, materialize)
        outs[fused] = (model.generate(ids, modal_inputs=feats, max_new_tokens=10, do_sample=False), model._dws.logits.clone(),
                       [k.clone() for k in model._dws.cache.k])
        assert getattr(model._dws, what.lower()) == fused
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    n = outs[True][0].shape[1]
    for a, b in zip(outs[True][2], outs[False][2]):
        L = model._dws.cache.length
        assert torch.equal(a[:, :, :L], b[:, :, :L])


def test_padded_text_only_batch_keeps_pads_masked(golden):
    """HF generate appends ones to the caller's mask: the padded prompt positions stay masked in every decode step."""
    dtype = torch.bfloat16
    model = d128_model(golden, dtype)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(3, 1000, (2, 12), generator=g).cuda()
    mask = torch.ones_like(ids)
    mask[1, :4] = 0  # left padding of the second prompt
    out = model.generate(ids, attention_mask=mask, max_new_tokens=4, do_sample=False)
    assert model._dws is not None and model._dws.key_mask is not None
    # row 1 alone, unpadded, must generate the same continuation
    solo = model.generate(ids[1:, 4:], max_new_tokens=4, do_sample=False)
    full = model.forward(out[1:, 4:], torch.ones_like(out[1:, 4:]))
    for i in range(4):
        lg = full.logits[0, 8 + i - 1].float()
        assert lg.max() - lg[out[1, 12 + i]] <= 2 ** -5 * lg.abs().max()
    assert solo.shape == (1, 12)


def test_full_width_decode_layer_vs_prefill():
    """vicuna-7B width (H 4096, I 11008, r 128, 32 heads), one decoder layer, 3-way composition: decode steps (stream-K skinny
    linears with the default group's rank 384 K-extension, split-KV attention) against the prefill of the extended sequence."""
    dtype = torch.bfloat16
    cfg, base, adapters = syn.make_composed_on_device(["video", "audio", "vision"], torch.device("cuda"), dtype, coeff=0.333, seed=4, layers=1)
    for mat in (False, True):
        model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfg), base, adapters, device="cuda", dtype=dtype, materialize=mat)
        B = 5
        g = torch.Generator().manual_seed(9)
        ids = syn.make_prompt_ids(B, ["vision", "audio"], 20, cfg["vocab_size"], 7, SO.MODAL_TOKEN_INDEXES, 6).cuda()
        feats = {"vision": torch.randn(B, 30, syn.MODAL_FEATURE_DIM["vision"], generator=g).to(dtype).cuda(),
                 "audio": torch.randn(B, 17, syn.MODAL_FEATURE_DIM["audio"], generator=g).to(dtype).cuda()}
        out_ids = model.generate(ids, modal_inputs=feats, max_new_tokens=4, do_sample=False)
        full = model.forward(out_ids, torch.ones_like(out_ids), modal_inputs=feats)
        o1 = model.forward(ids, torch.ones_like(ids), modal_inputs=feats, use_cache=True, cache_extra=8)
        cache = o1.past_key_values
        Sp = full.logits.shape[1]
        for i in range(3):
            tok = out_ids[:, ids.shape[1] + i:ids.shape[1] + i + 1]
            o = model.forward(tok, None, past_key_values=cache, modal_inputs=feats)
            logits_close(o.logits[:, 0], full.logits[:, Sp - 4 + i, :], "bf16", f"full-width decode step {i} (materialize={mat})")
        del model
