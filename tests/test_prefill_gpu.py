"""GPU parity of the composed-model prefill (splice -> projectors -> routed decoder layers -> logits).

Compared against (a) the fixtures the UNMODIFIED reference decoder layers produced (tests/golden/prefill_c1.pt) and
(b) the CPU oracle run on the same inputs.  Tolerance (bf16 tolerance stated per north_star; max-abs and cosine are
printed): logits / hidden max-abs <= 2^-5 (bf16) or 2^-8 (fp16) of the tensor's max magnitude and cosine >= 0.9995 /
0.99999 — the reference rounds after every op in 16 bits, the kernels accumulate in fp32 and round once per linear."""
import json
import os

import pytest
import torch

from modelcompose_b200 import builder as BD
from modelcompose_b200 import model as MD
from modelcompose_b200 import synthetic as syn
from oracle import merge_oracle as MO
from oracle import model_oracle as XO
from oracle import splice_oracle as SO

pytestmark = pytest.mark.gpu

STRATEGY_C1 = "online-merge-reset-default-vision=0.5,default-audio=0.5"
DTYPES = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}
MAXABS = {"torch.bfloat16": 2 ** -5, "torch.float16": 2 ** -8}
COS = {"torch.bfloat16": 0.9995, "torch.float16": 0.99999}


def compare(got, ref, key, what):
    got, ref = got.float().cpu().flatten(), ref.float().cpu().flatten()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    print(f"{what} [{key}]: max-abs {err:.4g} (scale {scale:.4g}, ratio {err / scale:.3g}) cosine {cos:.7f}")
    assert err <= MAXABS[key] * scale and cos >= COS[key], (what, err, scale, cos)


def tiny_model(golden, dtype):
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfg = MD.MultimodalConfig.from_dict(run["config"])
    base = syn.make_base_llm(seed=1)
    return MD.MultimodalLlamaForCausalLM(cfg, base, run["state_dict"], device="cuda", dtype=dtype), run, base


@pytest.mark.parametrize("key", list(DTYPES))
def test_decoder_layers_vs_reference_fixture(golden, key):
    dtype = DTYPES[key]
    model, run, base = tiny_model(golden, dtype)
    assert model.modal_names == ["default", "audio", "vision"]
    assert model.default_adapter_names == ["default-audio", "default-vision"]
    g = golden("prefill_c1.pt")
    ref = g["out"][key]
    x = g["x"].to(dtype).cuda()
    mid = torch.zeros(x.shape[:2], dtype=torch.uint8)
    for i, m in enumerate(model.modal_names):
        mid[g["masks"][m]] = i
    logits, _, hidden = model.prefill(x, mid.cuda(), None, output_hidden_states=True)
    torch.cuda.synchronize()
    compare(hidden[1], ref["hidden"][0], key, "hidden after layer 0")
    compare(hidden[2], ref["final_norm"], key, "final norm")
    compare(logits, ref["logits"], key, "logits")
    ws = next(iter(model._ws.values()))
    compare(ws.sequence_order(ws.x), ref["hidden"][1], key, "hidden after layer 1")
    # modal_id=None: every token takes the default adapter (decode-style path, multimodal_llama.py:436-438,:703-704)
    model.prefill(x, None, None)
    compare(ws.sequence_order(ws.x), ref["hidden_nomask"], key, "hidden (no modality mask)")


def oracle_forward(cfg: dict, base, sd, ids, attn, feats, dtype, n_heads):
    """CPU oracle of the whole path on 16-bit tensors: projector -> prefix/suffix -> splice -> routed layers -> logits."""
    modal_names = MD.infer_modals(MD.MultimodalConfig.from_dict(cfg))
    cpu = lambda t: t.detach().to("cpu", dtype)
    proj = {}
    for m in modal_names[1:]:
        if m not in feats:
            continue
        pre = f"model.modal_projectors.{m}."
        f = feats[m].reshape(feats[m].shape[0], -1, feats[m].shape[-1])
        proj[m] = XO.projector_forward(cpu(f), [cpu(sd[pre + "0.weight"]), cpu(sd[pre + "2.weight"])],
                                       [cpu(sd[pre + "0.bias"]), cpu(sd[pre + "2.bias"])])
    pre_t = {m: cpu(sd[f"prefix_tokens.{m}"]) for m in proj}
    suf_t = {m: cpu(sd[f"suffix_tokens.{m}"]) for m in proj}
    am, embeds, _, masks = SO.splice(ids, attn, None, cpu(base["model.embed_tokens.weight"]),
                                     SO.add_prefix_suffix(proj, pre_t, suf_t))
    names, scaling, dnames = MO.effective_scaling(modal_names, cfg["lora_r"], cfg["lora_alpha"], cfg["reset_scaling_weights"])
    layers = []
    for li in range(cfg["num_hidden_layers"]):
        layer = {"input_layernorm": cpu(base[f"model.layers.{li}.input_layernorm.weight"]),
                 "post_attention_layernorm": cpu(base[f"model.layers.{li}.post_attention_layernorm.weight"])}
        for ln in syn.LINEAR_NAMES:
            p = f"model.layers.{li}.{ln}."
            A = {k[len(p) + 7:-7]: cpu(v) for k, v in sd.items() if k.startswith(p + "lora_A.")}
            Bm = {k[len(p) + 7:-7]: cpu(v) for k, v in sd.items() if k.startswith(p + "lora_B.")}
            layer[ln.split(".")[1]] = XO.LinearParams(cpu(base[p + "weight"]), A, Bm, scaling, dnames)
        layers.append(layer)
    bmasks = {k: v.bool() for k, v in masks.items()}
    ordered = {m: bmasks.get(m, torch.zeros_like(bmasks["default"])) for m in modal_names}
    logits, _ = XO.model_forward(embeds, layers, cpu(base["model.norm.weight"]), cpu(base["lm_head.weight"]),
                                 ordered, modal_names, n_heads, cfg["rms_norm_eps"],
                                 rope_linear_factor=float((cfg.get("rope_scaling") or {}).get("factor", 1.0)))
    return logits, bmasks, modal_names


@pytest.mark.parametrize("key", list(DTYPES))
def test_end_to_end_forward_vs_oracle(golden, key):
    """input_ids with sentinels + encoder features -> projector -> splice -> routed layers -> logits, vs the CPU oracle
    evaluated in the same 16-bit dtype."""
    dtype = DTYPES[key]
    model, run, base = tiny_model(golden, dtype)
    g = torch.Generator().manual_seed(9)
    B, n_text = 3, 20
    ids = syn.make_prompt_ids(B, ["vision", "audio"], n_text, 1000, seed=3, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=6)
    feats = {"audio": torch.randn(B, 9, 48, generator=g).to(dtype), "vision": torch.randn(B, 14, 64, generator=g).to(dtype)}
    attn = torch.ones_like(ids)
    out = model.forward(ids.cuda(), attn.cuda(), modal_inputs={k: v.cuda() for k, v in feats.items()})
    torch.cuda.synchronize()
    logits, bmasks, _ = oracle_forward(run["config"], base, run["state_dict"], ids, attn, feats, dtype, 4)
    assert out.logits.shape == logits.shape
    assert torch.equal(out.modal_id.cpu() == 1, bmasks["audio"]) and torch.equal(out.modal_id.cpu() == 2, bmasks["vision"])
    compare(out.logits, logits, key, "end-to-end logits")


def test_rope_scaling_linear_vs_oracle_and_dynamic_refused(golden):
    """config.rope_scaling = {"type": "linear", "factor": 4} (multimodal_llama.py:193-199): the positions are divided by the factor;
    prefill vs the oracle, decode steps vs the prefill of the extended sequence.  "dynamic" is refused, unknown types get the reference's
    ValueError."""
    dtype, key = torch.bfloat16, "torch.bfloat16"
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    base = syn.make_base_llm(seed=1)
    cfgd = dict(run["config"])
    cfgd["rope_scaling"] = {"type": "linear", "factor": 4.0}
    model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfgd), base, run["state_dict"], device="cuda", dtype=dtype)
    plain = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(run["config"]), base, run["state_dict"], device="cuda", dtype=dtype)
    g = torch.Generator().manual_seed(4)
    B = 2
    ids = syn.make_prompt_ids(B, ["vision", "audio"], 24, 1000, seed=5, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=6)
    feats = {"audio": torch.randn(B, 9, 48, generator=g).to(dtype), "vision": torch.randn(B, 14, 64, generator=g).to(dtype)}
    attn = torch.ones_like(ids)
    dev_feats = {k: v.cuda() for k, v in feats.items()}
    out = model.forward(ids.cuda(), attn.cuda(), modal_inputs=dev_feats)
    logits, _, _ = oracle_forward(cfgd, base, run["state_dict"], ids, attn, feats, dtype, 4)
    compare(out.logits, logits, key, "logits with linear RoPE scaling")
    # the tables themselves: bit-identical to the oracle's scaled table, not the plain one; and the logits moved
    D, n = 64, 256
    want_cos, want_sin = XO.rope_cos_sin(D, n, dtype, linear_factor=4.0)
    assert torch.equal(model._rope[0][:n].cpu(), want_cos) and torch.equal(model._rope[1][:n].cpu(), want_sin)
    assert not torch.equal(model._rope[0][:n].cpu(), XO.rope_cos_sin(D, n, dtype)[0])
    unscaled = plain.forward(ids.cuda(), attn.cuda(), modal_inputs=dev_feats).logits
    assert not torch.equal(out.logits, unscaled)
    for bad, exc in (({"type": "dynamic", "factor": 2.0}, NotImplementedError), ({"type": "yarn", "factor": 2.0}, ValueError)):
        cfgd["rope_scaling"] = bad
        with pytest.raises(exc):
            MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfgd), base, run["state_dict"], device="cuda", dtype=dtype)


def test_modality_major_row_order_is_bit_identical(golden, monkeypatch):
    """the routed linears run on rows grouped by modality (one adapter group per tile); rows are independent and skipped
    LoRA blocks only ever multiply zeros, so the logits equal the sequence-order run bit for bit"""
    g = torch.Generator().manual_seed(19)
    B = 4
    ids = syn.make_prompt_ids(B, ["audio", "vision"], 25, 1000, seed=6, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=4)
    feats = {"audio": torch.randn(B, 30, 48, generator=g).to(torch.bfloat16).cuda(),
             "vision": torch.randn(B, 41, 64, generator=g).to(torch.bfloat16).cuda()}
    outs = []
    for flag in (True, False):
        monkeypatch.setattr(MD, "MODALITY_MAJOR", flag)
        model, _, _ = tiny_model(golden, torch.bfloat16)
        o = model.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats, output_hidden_states=True)
        ws = next(iter(model._ws.values()))
        assert ws.permute == flag and (flag is False or not ws.perm_is_identity)
        outs.append((o.logits.clone(), [h.clone() for h in o.hidden_states]))
    assert torch.equal(outs[0][0], outs[1][0])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][1], outs[1][1]))


@pytest.mark.parametrize("up_tuning", [3, 4])
def test_pair_kernel_prefill_is_bit_identical(golden, monkeypatch, up_tuning):
    """the base + LoRA-up launches on the CTA-pair kernels give the same logits as on the single-CTA kernel"""
    g = torch.Generator().manual_seed(23)
    B = 5
    ids = syn.make_prompt_ids(B, ["vision", "audio"], 40, 1000, seed=7, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=4)
    feats = {"audio": torch.randn(B, 50, 48, generator=g).to(torch.bfloat16).cuda(),
             "vision": torch.randn(B, 90, 64, generator=g).to(torch.bfloat16).cuda()}
    outs = []
    for t in (0, up_tuning):
        monkeypatch.setattr(MD, "UP_TUNING", t)
        model, _, _ = tiny_model(golden, torch.bfloat16)
        outs.append(model.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats).logits.clone())
    assert torch.equal(outs[0], outs[1])


def test_auto_variant_takes_the_pair_kernel_for_large_batches(golden, monkeypatch):
    """unset MC_LINEAR_UP_TUNING: batches of >= 8192 rows run their base + LoRA-up launches on the 512x256 CTA-pair kernel
    (smaller ones on the single-CTA kernel) — same logits bit for bit as the single-CTA kernel forced"""
    g = torch.Generator().manual_seed(29)
    B = 24
    ids = syn.make_prompt_ids(B, ["vision", "audio"], 40, 1000, seed=9, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=4)
    feats = {"audio": torch.randn(B, 60, 48, generator=g).to(torch.bfloat16).cuda(),
             "vision": torch.randn(B, 260, 64, generator=g).to(torch.bfloat16).cuda()}
    outs = []
    for t, want in ((None, 3), (0, 0)):
        monkeypatch.setattr(MD, "UP_TUNING", t)
        model, _, _ = tiny_model(golden, torch.bfloat16)
        outs.append(model.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats).logits.clone())
        ws = next(iter(model._ws.values()))
        assert ws.T >= MD.UP_TUNING_PAIR_MIN_ROWS and ws.up_tuning == want, (ws.T, ws.up_tuning)
    assert torch.equal(outs[0], outs[1])
    monkeypatch.setattr(MD, "UP_TUNING", None)
    model, _, _ = tiny_model(golden, torch.bfloat16)
    model.forward(ids[:2].cuda(), torch.ones_like(ids[:2]).cuda(), modal_inputs={k: v[:2] for k, v in feats.items()})
    assert next(iter(model._ws.values())).up_tuning == 0


@pytest.mark.parametrize("merged,coeff", [(["audio", "vision", "video"], 0.333), (["audio", "vision", "video", "point"], 0.25)])
def test_full_width_layer_vs_oracle(merged, coeff):
    """vicuna-7B WIDTH (H 4096, I 11008, r 128, vocab 32000, 5+5 prefix/suffix), one decoder layer, 2 requests with one block of
    every merged modality at reduced length (the CPU oracle runs the reference's dense-then-mask schedule, so the token count
    is kept where it finishes in seconds).  3 merged adapters at 0.333 (BASELINE configs 3 / 4: 4 routing groups, text rank 384)
    and 4 at 0.25 (config 5, MCUB-4: 5 routing groups, text rank 512, point-cloud projector with 384-wide features)."""
    dev = torch.device("cuda")
    cfg, base, sd = syn.make_composed_on_device(merged, dev, torch.bfloat16, coeff=coeff, seed=1, layers=1)
    model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfg), base, sd, device=dev, dtype=torch.bfloat16)
    assert model.modal_names == ["default"] + merged and model.rank_total == 128 * 2 * len(merged)
    g = torch.Generator().manual_seed(4)
    B = 2
    present = ["video", "vision", "audio"] + (["point"] if "point" in merged else [])
    ids = syn.make_prompt_ids(B, present, 30, cfg["vocab_size"], seed=5, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=10)
    feats = {"audio": torch.randn(B, 40, 768, generator=g).to(torch.bfloat16),
             "vision": torch.randn(B, 70, 1024, generator=g).to(torch.bfloat16),
             "video": torch.randn(B, 2, 45, 1024, generator=g).to(torch.bfloat16)}
    if "point" in merged:
        feats["point"] = torch.randn(B, 33, syn.MODAL_FEATURE_DIM["point"], generator=g).to(torch.bfloat16)
    attn = torch.ones_like(ids)
    out = model.forward(ids.cuda(), attn.cuda(), modal_inputs={k: v.cuda() for k, v in feats.items()})
    torch.cuda.synchronize()
    logits, bmasks, names = oracle_forward(cfg, base, sd, ids, attn, feats, torch.bfloat16, 32)
    assert out.logits.shape == logits.shape
    for i, m in enumerate(names):
        assert torch.equal(out.modal_id.cpu() == i, bmasks[m]), m
        assert int(bmasks[m].sum()) > 0, m   # every routing group is exercised
    compare(out.logits, logits, "torch.bfloat16", f"full-width one-layer logits, {len(names)} routing groups")


def test_materialised_weights_vs_oracle(golden):
    """W_eff,g built on the device (rank-r GEMM + merge-kernel blend) against the fp32 formula of the reference's own tooling
    (delta_weights_compare.py:24-31,61, restated in oracle/merge_oracle.materialise_effective_weight): within 1.5 ulp of the
    16-bit dtype relative to the weight's magnitude (the text group is a blend of ROUNDED dense checkpoints: two roundings)."""
    from modelcompose_b200 import materialize as MZ
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfg = MD.MultimodalConfig.from_dict(run["config"])
    base, sd = syn.make_base_llm(seed=1), run["state_dict"]
    names = MD.infer_modals(cfg)
    _, scaling, dnames = MD.adapter_scaling(names, cfg.lora_r, cfg.lora_alpha, cfg.reset_scaling_weights)
    for dtype, ulp in ((torch.bfloat16, 2.0 ** -8), (torch.float16, 2.0 ** -11)):
        for key in ("model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj"):
            W = base[key + ".weight"].to(dtype)
            A = {a: sd[f"{key}.lora_A.{a}.weight"].to(dtype) for a in names[1:] + dnames}
            Bm = {a: sd[f"{key}.lora_B.{a}.weight"].to(dtype) for a in A}
            got = MZ.effective_weights(W.cuda(), {a: t.cuda() for a, t in A.items()}, {a: t.cuda() for a, t in Bm.items()},
                                       scaling, names, dnames, cfg.lora_alpha / cfg.lora_r)
            torch.cuda.synchronize()
            assert len(got) == len(names)
            for gi, name in enumerate(names):
                members = dnames if gi == 0 else [name]
                want = MO.materialise_effective_weight(W, [A[a] for a in members], [Bm[a] for a in members],
                                                       [scaling[a] for a in members], out_dtype=torch.float32)
                err = (got[gi].float().cpu() - want).abs().max().item()
                assert err <= 1.5 * ulp * want.abs().max().item(), (dtype, key, name, err, want.abs().max().item())


@pytest.mark.parametrize("key", list(DTYPES))
def test_materialised_form_vs_reference_fixture_and_branch_form(golden, key):
    """The grouped-GEMM evaluation over materialised per-group weights against (a) the unmodified reference's decoder-layer
    outputs (same bar as the branch form: the dense form rounds W_eff once, the reference rounds every branch) and (b) this
    library's branch form on an end-to-end forward with every routing group present."""
    dtype = DTYPES[key]
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfg = MD.MultimodalConfig.from_dict(run["config"])
    base = syn.make_base_llm(seed=1)
    dense = MD.MultimodalLlamaForCausalLM(cfg, base, run["state_dict"], device="cuda", dtype=dtype, materialize=True)
    branch = MD.MultimodalLlamaForCausalLM(cfg, base, run["state_dict"], device="cuda", dtype=dtype, materialize=False)
    g = golden("prefill_c1.pt")
    ref = g["out"][key]
    x = g["x"].to(dtype).cuda()
    mid = torch.zeros(x.shape[:2], dtype=torch.uint8)
    for i, m in enumerate(dense.modal_names):
        mid[g["masks"][m]] = i
    logits, _, hidden = dense.prefill(x, mid.cuda(), None, output_hidden_states=True)
    torch.cuda.synchronize()
    compare(hidden[1], ref["hidden"][0], key, "materialised: hidden after layer 0")
    compare(logits, ref["logits"], key, "materialised: logits")
    dense.prefill(x, None, None)   # every row on the text group's weights
    ws = next(iter(dense._ws.values()))
    compare(ws.sequence_order(ws.x), ref["hidden_nomask"], key, "materialised: hidden (no modality mask)")
    gen = torch.Generator().manual_seed(31)
    B = 3
    ids = syn.make_prompt_ids(B, ["vision", "audio"], 20, 1000, seed=4, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=6)
    feats = {"audio": torch.randn(B, 9, 48, generator=gen).to(dtype).cuda(), "vision": torch.randn(B, 14, 64, generator=gen).to(dtype).cuda()}
    a = dense.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats).logits
    b = branch.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats).logits
    compare(a, b, key, "materialised vs branch form, end-to-end logits")


def test_materialised_full_width_layer_vs_oracle():
    """vicuna-7B width, one layer, 5 routing groups (MCUB-4), materialised form: every group's segment, the partial tiles at
    the segment boundaries and the CTA-pair kernel's 512-row tiles, against the CPU oracle (branch form)."""
    dev = torch.device("cuda")
    merged = ["audio", "vision", "video", "point"]
    cfg, base, sd = syn.make_composed_on_device(merged, dev, torch.bfloat16, coeff=0.25, seed=1, layers=1)
    model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfg), base, sd, device=dev, dtype=torch.bfloat16, materialize=True)
    g = torch.Generator().manual_seed(4)
    B = 2
    present = ["video", "vision", "audio", "point"]
    ids = syn.make_prompt_ids(B, present, 30, cfg["vocab_size"], seed=5, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=10)
    feats = {"audio": torch.randn(B, 40, 768, generator=g).to(torch.bfloat16), "vision": torch.randn(B, 70, 1024, generator=g).to(torch.bfloat16),
             "video": torch.randn(B, 2, 45, 1024, generator=g).to(torch.bfloat16),
             "point": torch.randn(B, 33, syn.MODAL_FEATURE_DIM["point"], generator=g).to(torch.bfloat16)}
    attn = torch.ones_like(ids)
    out = model.forward(ids.cuda(), attn.cuda(), modal_inputs={k: v.cuda() for k, v in feats.items()})
    torch.cuda.synchronize()
    logits, bmasks, names = oracle_forward(cfg, base, sd, ids, attn, feats, torch.bfloat16, 32)
    compare(out.logits, logits, "torch.bfloat16", "materialised full-width one-layer logits, 5 routing groups")
    # the pair kernel (512-row tiles) on the same rows gives the single-CTA kernel's bits
    import modelcompose_b200.model as MDm
    old = MDm.UP_TUNING
    try:
        MDm.UP_TUNING = 3
        model._ws.clear()
        out3 = model.forward(ids.cuda(), attn.cuda(), modal_inputs={k: v.cuda() for k, v in feats.items()})
        assert torch.equal(out3.logits, out.logits)
    finally:
        MDm.UP_TUNING = old


def test_loader_from_disk_and_text_only(tmp_path, golden):
    """merge CLI output dir + base dir -> load_pretrained_model (fp16 like the reference) -> forward."""
    from modelcompose_b200 import merge as MG
    (v_sd, v_cfg), (a_sd, a_cfg) = golden("merge_c1.pt")["inputs"]["vision"], golden("merge_c1.pt")["inputs"]["audio"]
    vdir, adir, odir, bdir = (str(tmp_path / n) for n in ("vision", "audio", "out-multimodal", "base"))
    syn.save_checkpoint_dir(vdir, v_sd, v_cfg)
    syn.save_checkpoint_dir(adir, a_sd, a_cfg)
    MG.merge_checkpoints([vdir, adir], odir, STRATEGY_C1)
    os.makedirs(bdir)
    torch.save(syn.make_base_llm(seed=1), os.path.join(bdir, "pytorch_model.bin"))
    tok, model, procs, ctx = BD.load_pretrained_model(odir, bdir, "out-multimodal")
    assert tok is None and procs is None and ctx == 2048 and model.dtype == torch.float16
    assert json.load(open(os.path.join(odir, "config.json")))["reset_scaling_weights"] == "default-vision=0.5,default-audio=0.5"
    ids = torch.randint(3, 1000, (2, 17), generator=torch.Generator().manual_seed(0)).cuda()
    out = model.forward(ids, torch.ones_like(ids))
    assert out.logits.shape == (2, 17, 1000) and torch.isfinite(out.logits).all()
    with pytest.raises(NotImplementedError):
        BD.load_pretrained_model(odir, bdir, "llava-thing")
    with pytest.raises(TypeError):
        model.forward(ids[:, :1], torch.ones_like(ids), past_key_values=[()])


def test_cli_writes_dense_merged_weights(tmp_path, golden):
    """merge CLI with --materialize-base: besides the reference's three output files, the reset-blended dense weights of the
    text path; checked against the fp32 formula, and the loader with materialize=True runs on the same directory."""
    from modelcompose_b200 import merge as MG
    (v_sd, v_cfg), (a_sd, a_cfg) = golden("merge_c1.pt")["inputs"]["vision"], golden("merge_c1.pt")["inputs"]["audio"]
    vdir, adir, odir, bdir = (str(tmp_path / n) for n in ("vision", "audio", "out-multimodal", "base"))
    syn.save_checkpoint_dir(vdir, v_sd, v_cfg)
    syn.save_checkpoint_dir(adir, a_sd, a_cfg)
    os.makedirs(bdir)
    base = syn.make_base_llm(seed=1)
    torch.save(base, os.path.join(bdir, "pytorch_model.bin"))
    MG.main([vdir, adir, "-o", odir, "--strategy", STRATEGY_C1, "--materialize-base", bdir])
    dense = torch.load(os.path.join(odir, "effective_weights.bin"))
    merged = torch.load(os.path.join(odir, "adapter_model.bin"))
    assert len(dense) == 2 * 7   # every decoder linear of the 2-layer model
    key = "model.layers.1.mlp.up_proj"
    got = dense[key + ".weight"]
    dt = base[key + ".weight"].dtype if base[key + ".weight"].dtype in (torch.float16, torch.bfloat16) else torch.float16
    assert got.dtype == dt   # the base checkpoint's 16-bit dtype (fp16 for an fp32 base, as the reference loads it)
    W = base[key + ".weight"].to(dt)
    members = ["default-vision", "default-audio"]
    want = MO.materialise_effective_weight(W, [merged[f"{key}.lora_A.{a}.weight"].to(dt) for a in members],
                                           [merged[f"{key}.lora_B.{a}.weight"].to(dt) for a in members],
                                           [16 / 8 * 0.5, 16 / 8 * 0.5], out_dtype=torch.float32)
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    assert (got.float() - want).abs().max().item() <= 1.5 * ulp * want.abs().max().item()
    tok, model, procs, ctx = BD.load_pretrained_model(odir, bdir, "out-multimodal", materialize=True)
    assert model.materialize and len(model.layers[0].Weff["q_proj"]) == 3
    ids = torch.randint(3, 1000, (2, 17), generator=torch.Generator().manual_seed(0)).cuda()
    assert torch.isfinite(model.forward(ids, torch.ones_like(ids)).logits).all()


@pytest.mark.parametrize("key", list(DTYPES))
def test_decode_step_and_generate(golden, key):
    """Decode steps (KV cache, default adapter only: multimodal_llama.py:436-438, multimodal_arch.py:290-293) must agree
    with a from-scratch prefill of the extended sequence, which is itself checked against the oracle above."""
    dtype = DTYPES[key]
    model, run, base = tiny_model(golden, dtype)
    g = torch.Generator().manual_seed(21)
    B = 3
    ids = syn.make_prompt_ids(B, ["vision", "audio"], 9, 1000, seed=8, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=4)
    feats = {"audio": torch.randn(B, 6, 48, generator=g).to(dtype).cuda(), "vision": torch.randn(B, 11, 64, generator=g).to(dtype).cuda()}
    new = 6
    out_ids = model.generate(ids.cuda(), modal_inputs=feats, max_new_tokens=new, do_sample=False)
    assert out_ids.shape == (B, ids.shape[1] + new) and torch.equal(out_ids[:, :ids.shape[1]].cpu(), ids)
    # teacher forcing: one prefill over prompt + generated tokens
    full = model.forward(out_ids, torch.ones_like(out_ids), modal_inputs=feats)
    Sp = full.logits.shape[1]
    for i in range(new):
        step_logits = full.logits[:, Sp - new + i - 1, :].float()
        tok = out_ids[:, ids.shape[1] + i]
        top = step_logits.max(-1).values
        chosen = step_logits.gather(1, tok[:, None]).squeeze(1)
        assert ((top - chosen) <= MAXABS[key] * step_logits.abs().max()).all(), (i, top, chosen)
    # explicit decode-step logits vs the prefill logits at the same position
    o1 = model.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats, use_cache=True)
    cache = o1.past_key_values
    assert isinstance(cache, MD.KVCache) and cache.length == o1.logits.shape[1] and len(cache.legacy()) == 2
    t1 = out_ids[:, ids.shape[1]:ids.shape[1] + 1]
    o2 = model.forward(t1, torch.ones((B, cache.length + 1), dtype=torch.int64, device="cuda"), past_key_values=cache, modal_inputs=feats)
    assert cache.length == o1.logits.shape[1] + 1 and o2.logits.shape == (B, 1, 1000)
    compare(o2.logits[:, 0], full.logits[:, Sp - new, :], key, "decode-step logits vs prefill of the extended sequence")
    # generate()'s prefill computes the lm_head on the last position only: bit-identical to that row of the full logits
    o3 = model.forward(ids.cuda(), torch.ones_like(ids).cuda(), modal_inputs=feats, last_logits_only=True)
    assert o3.logits.shape == (B, 1, 1000) and torch.equal(o3.logits[:, 0], o1.logits[:, -1])
    # sampling path runs and respects eos
    s_ids = model.generate(ids.cuda(), modal_inputs=feats, max_new_tokens=4, do_sample=True, temperature=0.7, top_p=0.9,
                           generator=torch.Generator(device="cuda").manual_seed(0), eos_token_id=int(out_ids[0, ids.shape[1]]))
    assert s_ids.shape[0] == B and ids.shape[1] < s_ids.shape[1] <= ids.shape[1] + 4
