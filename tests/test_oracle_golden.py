"""Pin oracle/ to the fixtures produced by the real reference (tests/golden/make_golden.py).

Integer / copy / elementwise paths must be bit-exact.  GEMM-bearing paths are compared with a
tolerance because CPU matmul kernels differ between hosts (AMX / AVX512-bf16 vs emulation):
fp32 1e-5 relative, bf16/fp16 2 ulp of the output scale.
"""
import json

import pytest
import torch

from oracle import merge_oracle as MO
from oracle import model_oracle as XO
from oracle import splice_oracle as SO

STRATEGY_C1 = "online-merge-reset-default-vision=0.5,default-audio=0.5"
DTYPES = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}
TOL = {"torch.float32": 2e-5, "torch.bfloat16": 2 ** -6, "torch.float16": 2 ** -9}


def _close(a, b, key):
    a, b = a.float(), b.float()
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err <= TOL[key] * scale, (key, err, scale)


# ------------------------------------------------------------------------------------------- merge (A1-A5)
@pytest.mark.parametrize("strategy", [STRATEGY_C1, "sum", "mean"])
def test_merge_weights_bit_exact(golden, strategy):
    g = golden("merge_c1.pt")
    (v_sd, v_cfg), (a_sd, a_cfg) = g["inputs"]["vision"], g["inputs"]["audio"]
    run = g["runs"][strategy]
    merged = MO.merge_weights([v_sd, a_sd], [v_cfg, a_cfg], strategy)
    assert list(merged.keys()) == list(run["state_dict"].keys())
    for k in merged:
        assert merged[k].dtype == run["state_dict"][k].dtype
        assert torch.equal(merged[k].view(torch.int16), run["state_dict"][k].view(torch.int16)), k


@pytest.mark.parametrize("strategy", [STRATEGY_C1, "sum", "mean"])
def test_merge_config_and_info(golden, strategy):
    g = golden("merge_c1.pt")
    cfgs = [g["inputs"]["vision"][1], g["inputs"]["audio"][1]]
    run = g["runs"][strategy]
    mcfg, after = MO.merge_configs(cfgs, strategy)
    assert mcfg == run["config"] and list(mcfg.keys()) == list(run["config"].keys())
    assert json.dumps(mcfg, indent=4) == run["config_json_text"]
    assert MO.merge_info_text(["{IN0}", "{IN1}"], after, "{OUT}") == run["merge_info"]


def test_get_modal_from_config():
    assert MO.get_modal_from_config({"mm_vision_tower": "x"}) == "vision"
    assert MO.get_modal_from_config({"mm_vision_tower": "", "mm_audio_encoder": "b"}) == "audio"
    assert MO.get_modal_from_config({"mm_video_encoder": "v", "mm_audio_encoder": "b"}) == "video"
    with pytest.raises(AssertionError):
        MO.get_modal_from_config({"mm_vision_tower": None})


def test_weighted_merge_matches_ref_sum_when_exact():
    # with weights 1.0 and values on a coarse grid every partial sum is exact in bf16 → both orders agree
    g = torch.Generator().manual_seed(0)
    ts = [(torch.randint(-8, 8, (1000,), generator=g).float() / 4).to(torch.bfloat16) for _ in range(3)]
    assert torch.equal(MO.weighted_merge(ts, [1, 1, 1]), MO.ref_sum(ts))


# ------------------------------------------------------------------------------------------- linear (A8, A9)
def test_effective_scaling(golden):
    g = golden("linear_c1.pt")
    case = g["cases"]["self_attn.q_proj"]
    names, scaling, dnames = MO.effective_scaling(g["modal_names"], 8, 16, "default-vision=0.5,default-audio=0.5")
    assert names == case["adapters"] and dnames == case["default_adapter_names"]
    assert scaling == case["scaling"]
    # 7B coefficients: 2.0 * 0.333 in float64
    _, s7, _ = MO.effective_scaling(["default", "audio", "vision", "video"], 128, 256,
                                    "default-video=0.333,default-audio=0.333,default-vision=0.333")
    assert s7["default-video"] == 2.0 * 0.333 and s7["video"] == 2.0


def _linear_params(golden, lname, dt):
    from modelcompose_b200 import synthetic as syn
    g = golden("linear_c1.pt")
    m = golden("merge_c1.pt")["runs"][STRATEGY_C1]["state_dict"]
    case = g["cases"][lname]
    base = syn.make_base_llm(seed=1)
    pre = f"model.layers.0.{lname}."
    A = {k.split(".")[1] if False else k[len(pre) + 7:-7]: v.to(dt) for k, v in m.items() if k.startswith(pre + "lora_A.")}
    B = {k[len(pre) + 7:-7]: v.to(dt) for k, v in m.items() if k.startswith(pre + "lora_B.")}
    return case, base[pre + "weight"].to(dt), A, B


@pytest.mark.parametrize("lname", ["self_attn.q_proj", "mlp.down_proj"])
@pytest.mark.parametrize("key", list(DTYPES))
def test_lora_linear_forward(golden, lname, key):
    dt = DTYPES[key]
    case, W, A, B = _linear_params(golden, lname, dt)
    names = golden("linear_c1.pt")["modal_names"]
    out = XO.lora_linear_forward(case["x"].to(dt), W, A, B, case["scaling"], names, case["default_adapter_names"])
    assert list(out.keys()) == [k for k in case["out"][key] if k != "__base__"]
    for k in out:
        _close(out[k], case["out"][key][k], key)
    _close(XO.lora_linear_forward(case["x"].to(dt), W, A, B, case["scaling"], None), case["out"][key]["__base__"], key)
    # materialised W_eff reproduces the 'default' branch (reference's own formula, delta_weights_compare.py:24-31,61)
    if dt == torch.float32:
        dn = case["default_adapter_names"]
        Weff = MO.materialise_effective_weight(W, [A[n] for n in dn], [B[n] for n in dn], [case["scaling"][n] for n in dn])
        _close(torch.nn.functional.linear(case["x"], Weff), case["out"][key]["default"], key)


# ------------------------------------------------------------------------------------------- projector (A10)
@pytest.mark.parametrize("modal", ["vision", "audio"])
@pytest.mark.parametrize("key", list(DTYPES))
def test_projector(golden, modal, key):
    dt = DTYPES[key]
    g = golden("projector.pt")[modal]
    m = golden("merge_c1.pt")["runs"][STRATEGY_C1]["state_dict"]
    pre = f"model.modal_projectors.{modal}."
    y = XO.projector_forward(g["x"].to(dt), [m[pre + "0.weight"].to(dt), m[pre + "2.weight"].to(dt)],
                             [m[pre + "0.bias"].to(dt), m[pre + "2.bias"].to(dt)])
    _close(y, g[key], key)


def test_projector_linear(golden):
    g = golden("projector.pt")["linear"]
    _close(XO.projector_forward(g["x"], [g["weight"]], [g["bias"]]), g["torch.float32"], "torch.float32")


# ------------------------------------------------------------------------------------------- splice (A11-A13)
def _run_splice_oracle(case):
    feats = SO.add_prefix_suffix({m: case["features"][m] for m in case["modals"]}, case["prefix"], case["suffix"])
    return SO.splice(case["input_ids"], case["attention_mask"], case["labels"], case["embed"], feats,
                     case["modal_inputs_keys"])


def test_splice_cases_bit_exact(golden):
    cases = golden("splice.pt")
    assert len(cases) == 7
    for case in cases:
        if "raises" in case:
            with pytest.raises(Exception) as ei:
                _run_splice_oracle(case)
            assert type(ei.value).__name__ == case["raises"], case["name"]
            continue
        attn, embeds, labels, masks = _run_splice_oracle(case)
        ref = case["out"]
        assert torch.equal(embeds, ref["inputs_embeds"]), case["name"]
        assert attn.dtype == ref["attention_mask"].dtype and torch.equal(attn, ref["attention_mask"]), case["name"]
        if ref["labels"] is None:
            assert labels is None
        else:
            assert torch.equal(labels, ref["labels"]), case["name"]
        assert list(masks.keys()) == list(ref["modal_attention_mask"].keys()), case["name"]
        for k in masks:
            assert masks[k].dtype == ref["modal_attention_mask"][k].dtype, (case["name"], k)
            assert torch.equal(masks[k], ref["modal_attention_mask"][k]), (case["name"], k)


# ------------------------------------------------------------------------------------------- prefill (A14-A16)
def build_oracle_layers(golden, dt):
    from modelcompose_b200 import synthetic as syn
    m = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    base = syn.make_base_llm(seed=1)
    names, scaling, dnames = MO.effective_scaling(["default", "audio", "vision"], 8, 16, m["config"]["reset_scaling_weights"])
    layers = []
    for li in range(2):
        layer = {"input_layernorm": base[f"model.layers.{li}.input_layernorm.weight"].to(dt),
                 "post_attention_layernorm": base[f"model.layers.{li}.post_attention_layernorm.weight"].to(dt)}
        for ln in syn.LINEAR_NAMES:
            pre = f"model.layers.{li}.{ln}."
            A = {k[len(pre) + 7:-7]: v.to(dt) for k, v in m["state_dict"].items() if k.startswith(pre + "lora_A.")}
            B = {k[len(pre) + 7:-7]: v.to(dt) for k, v in m["state_dict"].items() if k.startswith(pre + "lora_B.")}
            layer[ln.split(".")[1]] = XO.LinearParams(base[pre + "weight"].to(dt), A, B, scaling, dnames)
        layers.append(layer)
    return layers, base


@pytest.mark.parametrize("key", list(DTYPES))
def test_prefill_layers_and_logits(golden, key):
    from tests.golden.make_golden import sd_digest
    dt = DTYPES[key]
    g = golden("prefill_c1.pt")
    layers, base = build_oracle_layers(golden, dt)
    assert sd_digest(base) == g["base_digest"], "synthetic base LLM generator drifted from the fixtures"
    ref = g["out"][key]
    x = g["x"].to(dt)
    S = x.shape[1]
    pos = torch.arange(S)[None]
    add = XO.causal_additive_mask(x.shape[0], S, dt)
    h = x
    for li, layer in enumerate(layers):
        h = XO.decoder_layer_forward(h, layer, g["masks"], g["modal_names"], 4, pos, add, 1e-5)
        _close(h, ref["hidden"][li], key)
    logits, hn = XO.model_forward(x, layers, base["model.norm.weight"].to(dt), base["lm_head.weight"].to(dt),
                                  g["masks"], g["modal_names"], 4, 1e-5)
    _close(hn, ref["final_norm"], key)
    _close(logits, ref["logits"], key)
    h0 = x
    for layer in layers:
        h0 = XO.decoder_layer_forward(h0, layer, None, g["modal_names"], 4, pos, add, 1e-5)
    _close(h0, ref["hidden_nomask"], key)


# ------------------------------------------------------------------------------------------- rope_scaling "linear"
@pytest.mark.parametrize("factor", [2.0, 4.0, 3.0])
def test_rope_linear_scaling_vs_transformers(factor):
    """LlamaLinearScalingRotaryEmbedding (multimodal_llama.py:193-199 -> transformers 4.31, a pinned dependency of the reference that is
    neither vendored under /root/reference nor installed here) divides the positions by the factor.  Anchor: the installed
    transformers' "linear" rope initialisation, which scales inv_freq instead — the same table bit for bit when the factor is a
    power of two ((t / f) * w == t * (w / f) in fp32), and up to the rounding of a ~10^3 rad argument otherwise."""
    from transformers import LlamaConfig
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    D, n = 128, 4096
    cfg = LlamaConfig(hidden_size=4 * D, num_attention_heads=4, max_position_embeddings=n, rope_scaling={"rope_type": "linear", "factor": factor})
    inv_freq, attention_scale = ROPE_INIT_FUNCTIONS["linear"](cfg, "cpu")
    assert attention_scale == 1.0
    freqs = torch.einsum("i,j->ij", torch.arange(n, dtype=torch.float32), inv_freq.float())
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = XO.rope_cos_sin(D, n, torch.float32, linear_factor=factor)
    if factor in (2.0, 4.0):
        assert torch.equal(cos, emb.cos()) and torch.equal(sin, emb.sin())
    else:
        assert (cos - emb.cos()).abs().max() < 1e-3 and (sin - emb.sin()).abs().max() < 1e-3
    plain = XO.rope_cos_sin(D, n, torch.float32)[0]
    assert torch.equal(cos[::int(factor)][: n // 4], plain[: n // 4]) if factor in (2.0, 4.0) else not torch.equal(cos, plain)
