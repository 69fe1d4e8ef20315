"""Deterministic synthetic checkpoints and inputs.

There is no network: every benchmark/test input is seeded random data of the shapes the
reference's training scripts produce (``train_multimodal.py:516-521`` saves every
``requires_grad`` parameter as ``adapter_model.bin``).  Used by tests, ``bench.py`` and
``__graft_entry__.smoke()``; contains no merge/forward arithmetic.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Tuple

import torch

# (name, in_features_attr, out_features_attr) of the 7 LocalLoraLinear per decoder layer
# (multimodal_llama.py:184-187, :357-359)
LINEAR_NAMES = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"]

# which config key announces each modality's encoder (merge_unimodal_modelcompose.py:15-21)
MODAL_ENCODER_KEY = {"vision": "mm_vision_tower", "audio": "mm_audio_encoder", "video": "mm_video_encoder",
                     "point": "mm_point_encoder"}
# projector config keys read by multimodal_projector/builder.py:203-244
MODAL_HIDDEN_KEY = {"vision": "mm_hidden_size", "audio": "mm_audio_hidden_size", "video": "mm_video_hidden_size",
                    "point": "mm_point_hidden_size"}
MODAL_PROJ_KEY = {"vision": "mm_projector_type", "audio": "mm_audio_projector_type",
                  "video": "mm_video_projector_type", "point": "mm_point_projector_type"}

TINY = dict(hidden_size=256, intermediate_size=688, num_attention_heads=4, num_key_value_heads=4,
            num_hidden_layers=2, vocab_size=1000, rms_norm_eps=1e-5, max_position_embeddings=2048,
            hidden_act="silu")
VICUNA_7B = dict(hidden_size=4096, intermediate_size=11008, num_attention_heads=32, num_key_value_heads=32,
                 num_hidden_layers=32, vocab_size=32000, rms_norm_eps=1e-5, max_position_embeddings=4096,
                 hidden_act="silu")
# encoder feature dims / token counts (CLIP-L 576x1024, BEATs 256x768 synthetic, LanguageBind 8x257x1024, PointBERT 513x384)
MODAL_FEATURE_DIM = {"vision": 1024, "audio": 768, "video": 1024, "point": 384}
MODAL_TOKENS = {"vision": 576, "audio": 256, "video": 2056, "point": 513}


def linear_shape(llama: dict, name: str) -> Tuple[int, int]:
    """(out_features, in_features) of a decoder linear."""
    H, I = llama["hidden_size"], llama["intermediate_size"]
    if name.endswith(("gate_proj", "up_proj")):
        return I, H
    if name.endswith("down_proj"):
        return H, I
    return H, H


def _uniform(gen, shape, bound, dtype):
    return ((torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound).to(dtype)


def _normal(gen, shape, std, dtype):
    return (torch.randn(shape, generator=gen, dtype=torch.float32) * std).to(dtype)


def make_unimodal_checkpoint(modal: str, seed: int, llama: dict = TINY, r: int = 8, lora_alpha: int = 16,
                             feat_dim: int = 64, n_prefix: int = 5, n_suffix: int = 5,
                             dtype=torch.bfloat16, projector_type: str = "mlp2x_gelu",
                             layers: int | None = None) -> Tuple[Dict[str, torch.Tensor], dict]:
    """One DAMC unimodal checkpoint (``lora_strategy modal+language``): adapters ``default`` and
    ``{modal}`` on all 7 linears of every layer, the modality projector, prefix/suffix tokens.
    Key order follows ``named_parameters()`` of the reference model.  LoRA A ~ U(±1/sqrt(in)) (peft
    init), LoRA B ~ N(0, 0.02) (peft's zeros would make parity vacuous), projector = nn.Linear default init."""
    g = torch.Generator().manual_seed(seed)
    H = llama["hidden_size"]
    L = llama["num_hidden_layers"] if layers is None else layers
    sd: Dict[str, torch.Tensor] = {}
    if n_prefix:
        for a in ("default", modal):
            sd[f"prefix_tokens.{a}"] = _normal(g, (1, n_prefix, H), 0.02, dtype)
    if n_suffix:
        for a in ("default", modal):
            sd[f"suffix_tokens.{a}"] = _normal(g, (1, n_suffix, H), 0.02, dtype)
    if projector_type == "linear":
        dims = [(H, feat_dim)]
        idx = [None]
    else:
        depth = int(projector_type[3:projector_type.index("x")])
        dims = [(H, feat_dim)] + [(H, H)] * (depth - 1)
        idx = [2 * i for i in range(depth)]
    for (o, i), k in zip(dims, idx):
        stem = f"model.modal_projectors.{modal}" + ("" if k is None else f".{k}")
        sd[f"{stem}.weight"] = _uniform(g, (o, i), 1.0 / math.sqrt(i), dtype)
        sd[f"{stem}.bias"] = _uniform(g, (o,), 1.0 / math.sqrt(i), dtype)
    for l in range(L):
        for name in LINEAR_NAMES:
            out_f, in_f = linear_shape(llama, name)
            for a in ("default", modal):
                sd[f"model.layers.{l}.{name}.lora_A.{a}.weight"] = _uniform(g, (r, in_f), 1.0 / math.sqrt(in_f), dtype)
            for a in ("default", modal):
                sd[f"model.layers.{l}.{name}.lora_B.{a}.weight"] = _normal(g, (out_f, r), 0.02, dtype)
    cfg = dict(llama)
    cfg.update({"model_type": "multimodal", "architectures": ["MultimodalLlamaForCausalLM"],
                "lora_strategy": "modal+language", "lora_r": r, "lora_alpha": lora_alpha, "lora_dropout": 0.05,
                "local_prefix_tokens": n_prefix, "local_suffix_tokens": n_suffix,
                MODAL_ENCODER_KEY[modal]: f"synthetic-{modal}-encoder",
                MODAL_HIDDEN_KEY[modal]: feat_dim, MODAL_PROJ_KEY[modal]: projector_type})
    if modal == "vision":
        cfg["mm_vision_encoder"] = cfg["mm_vision_tower"]
    return sd, cfg


def same_strategy_checkpoint(modal: str, seed: int, feat_dim: int, layers: int = 1) -> Tuple[Dict[str, torch.Tensor], dict]:
    """A ``--lora_strategy same`` checkpoint (reference train_multimodal.py:448-452): one ``default`` adapter on every
    linear plus the projector — what the ``convert-`` strategies of the merge CLI take (merge_unimodal_modelcompose.py:42-71)."""
    sd, cfg = make_unimodal_checkpoint(modal, seed=seed, feat_dim=feat_dim, layers=layers, n_prefix=0, n_suffix=0)
    sd = {k: v for k, v in sd.items() if f".{modal}.weight" not in k}
    return sd, dict(cfg, lora_strategy="same", num_hidden_layers=layers)


def ties_cli_checkpoints():
    """Inputs of the ties-* / convert-* CLI fixtures (tests/golden/ties.pt): two 1-layer DAMC checkpoints and two
    1-layer ``same``-strategy checkpoints, (state_dict, config) per modality."""
    damc = {}
    for m, s, f in (("vision", 110, 32), ("audio", 111, 24)):
        sd, cfg = make_unimodal_checkpoint(m, seed=s, feat_dim=f, layers=1)
        damc[m] = (sd, dict(cfg, num_hidden_layers=1))
    same = {m: same_strategy_checkpoint(m, s, f) for m, s, f in (("vision", 120, 32), ("audio", 121, 24))}
    return damc, same


def save_checkpoint_dir(path: str, sd: Dict[str, torch.Tensor], cfg: dict) -> None:
    os.makedirs(path, exist_ok=True)
    torch.save(sd, os.path.join(path, "adapter_model.bin"))
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=4)


def make_base_llm(seed: int, llama: dict = TINY, dtype=torch.bfloat16, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random-init base LLM state dict with vicuna key names (what ``model_base`` provides)."""
    g = torch.Generator().manual_seed(seed)
    H, V = llama["hidden_size"], llama["vocab_size"]
    sd = {"model.embed_tokens.weight": _normal(g, (V, H), std, dtype)}
    for l in range(llama["num_hidden_layers"]):
        for name in LINEAR_NAMES:
            sd[f"model.layers.{l}.{name}.weight"] = _normal(g, linear_shape(llama, name), std, dtype)
        sd[f"model.layers.{l}.input_layernorm.weight"] = (1.0 + _normal(g, (H,), 0.02, torch.float32)).to(dtype)
        sd[f"model.layers.{l}.post_attention_layernorm.weight"] = (1.0 + _normal(g, (H,), 0.02, torch.float32)).to(dtype)
    sd["model.norm.weight"] = (1.0 + _normal(g, (H,), 0.02, torch.float32)).to(dtype)
    sd["lm_head.weight"] = _normal(g, (V, H), std, dtype)
    return sd


def dense_7b_tensor_shapes(llama: dict = VICUNA_7B) -> List[Tuple[str, Tuple[int, ...]]]:
    """The 291 dense tensors of a vicuna-7B-shaped checkpoint (6,738,415,616 elements) — config C2."""
    H, V = llama["hidden_size"], llama["vocab_size"]
    shapes = [("model.embed_tokens.weight", (V, H))]
    for l in range(llama["num_hidden_layers"]):
        for name in LINEAR_NAMES:
            shapes.append((f"model.layers.{l}.{name}.weight", linear_shape(llama, name)))
        shapes.append((f"model.layers.{l}.input_layernorm.weight", (H,)))
        shapes.append((f"model.layers.{l}.post_attention_layernorm.weight", (H,)))
    shapes.append(("model.norm.weight", (H,)))
    shapes.append(("lm_head.weight", (V, H)))
    return shapes


def shard_tensors_greedy(sizes: List[int], world: int) -> List[List[int]]:
    """Greedy largest-first size balancing of tensor indices over ``world`` ranks; no collective."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += sizes[i]
    return [sorted(x) for x in out]


def make_prompt_ids(batch: int, modal_order: List[str], n_text: int, vocab: int, seed: int,
                    modal_token_indexes: Dict[str, int], n_head: int = 36) -> torch.Tensor:
    """Equal-length prompts: ``n_head`` system/USER tokens, then one sentinel per modality each followed by
    2 separator tokens, then ``n_text`` question tokens (ragged inference batches crash the reference,
    multimodal_arch.py:414-429, so every request shares the layout).  Token ids uniform in [3, vocab)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for _ in range(batch):
        parts = [torch.randint(3, vocab, (n_head,), generator=g)]
        for m in modal_order:
            parts.append(torch.tensor([modal_token_indexes[m]]))
            parts.append(torch.randint(3, vocab, (2,), generator=g))
        parts.append(torch.randint(3, vocab, (n_text,), generator=g))
        rows.append(torch.cat(parts))
    return torch.stack(rows).to(torch.int64)


def make_composed_on_device(modals: List[str], device, dtype=torch.bfloat16, llama: dict = VICUNA_7B, r: int = 128,
                            lora_alpha: int = 256, coeff: float = 0.333, n_prefix: int = 5, n_suffix: int = 5,
                            seed: int = 1, layers: int | None = None, feature_dims: Dict[str, int] | None = None):
    """A composed (online-merge-reset) model generated directly in device memory: returns
    ``(config dict, base_state_dict, merged_adapter_state_dict)`` with the key schema the reference merge CLI writes
    (``lora_{A,B}.default-{modal}`` + ``lora_{A,B}.{modal}`` per linear, projectors, prefix/suffix tokens;
    merge_unimodal_modelcompose.py:94-103).  Same distributions as ``make_base_llm`` / ``make_unimodal_checkpoint``;
    used by bench.py and the full-size GPU tests where CPU generation of 7B tensors would dominate the run."""
    g = torch.Generator(device=device).manual_seed(seed)
    cfg = dict(llama)
    if layers is not None:
        cfg["num_hidden_layers"] = layers
    H, V, L = cfg["hidden_size"], cfg["vocab_size"], cfg["num_hidden_layers"]
    fd = dict(MODAL_FEATURE_DIM)
    fd.update(feature_dims or {})

    def normal(shape, std):
        return torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, std, generator=g).to(dtype)

    def uniform(shape, bound):
        return torch.empty(shape, dtype=torch.float32, device=device).uniform_(-bound, bound, generator=g).to(dtype)

    base = {"model.embed_tokens.weight": normal((V, H), 0.02)}
    ad: Dict[str, torch.Tensor] = {}
    for m in modals:
        d = fd[m]
        ad[f"model.modal_projectors.{m}.0.weight"] = uniform((H, d), 1.0 / math.sqrt(d))
        ad[f"model.modal_projectors.{m}.0.bias"] = uniform((H,), 1.0 / math.sqrt(d))
        ad[f"model.modal_projectors.{m}.2.weight"] = uniform((H, H), 1.0 / math.sqrt(H))
        ad[f"model.modal_projectors.{m}.2.bias"] = uniform((H,), 1.0 / math.sqrt(H))
        if n_prefix:
            ad[f"prefix_tokens.{m}"] = normal((1, n_prefix, H), 0.02)
        if n_suffix:
            ad[f"suffix_tokens.{m}"] = normal((1, n_suffix, H), 0.02)
    for l in range(L):
        for name in LINEAR_NAMES:
            out_f, in_f = linear_shape(cfg, name)
            base[f"model.layers.{l}.{name}.weight"] = normal((out_f, in_f), 0.02)
            for m in modals:
                for a in (f"default-{m}", m):
                    ad[f"model.layers.{l}.{name}.lora_A.{a}.weight"] = uniform((r, in_f), 1.0 / math.sqrt(in_f))
                    ad[f"model.layers.{l}.{name}.lora_B.{a}.weight"] = normal((out_f, r), 0.02)
        base[f"model.layers.{l}.input_layernorm.weight"] = (1.0 + normal((H,), 0.02).float()).to(dtype)
        base[f"model.layers.{l}.post_attention_layernorm.weight"] = (1.0 + normal((H,), 0.02).float()).to(dtype)
    base["model.norm.weight"] = (1.0 + normal((H,), 0.02).float()).to(dtype)
    base["lm_head.weight"] = normal((V, H), 0.02)
    cfg.update({"model_type": "multimodal", "architectures": ["MultimodalLlamaForCausalLM"],
                "lora_strategy": "modal+language", "lora_r": r, "lora_alpha": lora_alpha, "lora_dropout": 0.05,
                "local_prefix_tokens": n_prefix, "local_suffix_tokens": n_suffix,
                "reset_scaling_weights": ",".join(f"default-{m}={coeff}" for m in modals)})
    for m in modals:
        cfg[MODAL_ENCODER_KEY[m]] = f"synthetic-{m}-encoder"
        cfg[MODAL_HIDDEN_KEY[m]] = fd[m]
        cfg[MODAL_PROJ_KEY[m]] = "mlp2x_gelu"
        if m == "vision":
            cfg["mm_vision_encoder"] = cfg["mm_vision_tower"]
    return cfg, base, ad


def prefill_flops(cfg: dict, modals_present: Dict[str, int], n_text: int, n_modal_adapters: int, r: int,
                  feature_dims: Dict[str, int] | None = None, n_prefix: int = 5, n_suffix: int = 5) -> Dict[str, float]:
    """ALGORITHMIC FLOPs of ONE sequence: every token runs the 7 base linears per layer; a modality
    token (features + prefix/suffix rows) adds one rank-r adapter, a text token the N concatenated default sub-adapters;
    projector 2·(d·H + H·H) per feature row; lm_head on every position; causal attention 2·2·S²·H/2 per layer."""
    H, I, V, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["num_hidden_layers"]
    fd = dict(MODAL_FEATURE_DIM)
    fd.update(feature_dims or {})
    per_tok_base = 2.0 * L * (4 * H * H + 3 * H * I)
    per_tok_lora = 2.0 * L * r * (4 * 2 * H + 3 * (H + I))
    n_modal_rows = sum(n + n_prefix + n_suffix for n in modals_present.values())
    S = n_text + n_modal_rows
    out = {"seq_len": S,
           "base": per_tok_base * S,
           "lora": per_tok_lora * (n_modal_rows + n_modal_adapters * n_text),
           "projector": sum(2.0 * n * (fd[m] * H + H * H) for m, n in modals_present.items()),
           "lm_head": 2.0 * V * H * S,
           "attention": L * 2.0 * S * S * H}
    out["linears"] = out["base"] + out["lora"] + out["projector"] + out["lm_head"]
    out["total"] = out["linears"] + out["attention"]
    return out
