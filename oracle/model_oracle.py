"""Oracle for the composed-model prefill.  TEST INFRASTRUCTURE.

torch-CPU restatement of ``modelcompose/model/language_model/multimodal_llama.py`` (line numbers
into that file unless another is named).  Every op runs in the tensor's own dtype in the same
order as the reference, so rounding points match the reference's eager path.  Pinned against fixtures produced by
running the reference itself (tests/golden/prefill_c1.pt, tests/test_oracle_golden.py); the one part no reference artefact
pins is ``linear_factor`` of ``rope_cos_sin`` (transformers 4.31's LlamaLinearScalingRotaryEmbedding, a dependency absent
here): it is anchored on the installed transformers' "linear" rope initialisation instead.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- A10 projector
def projector_forward(x: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> torch.Tensor:
    """multimodal_projector/builder.py:202-219 — ``linear`` (1 layer) or ``mlp{N}x_gelu``:
    Linear → [GELU(erf) → Linear]*(N-1)."""
    h = F.linear(x, weights[0], biases[0])
    for w, b in zip(weights[1:], biases[1:]):
        h = F.linear(F.gelu(h), w, b)
    return h


# ----------------------------------------------------------------------------- A9 LocalLoraLinear
def lora_linear_forward(x, W, lora_A: Dict[str, torch.Tensor], lora_B: Dict[str, torch.Tensor],
                        scaling: Dict[str, float], active_adapters=None,
                        default_adapter_names: Optional[Sequence[str]] = None, bias=None):
    """:120-160 (eval mode: dropout is identity).  Returns dict name→tensor, or the bare base
    output when ``active_adapters`` is falsy."""
    prev = x.dtype
    base = F.linear(x, W, bias)
    if not active_adapters:
        return base
    out = {}
    for name in active_adapters:
        # The reference constructs an adapter for every modal name (:84-88), so its ``not in self.lora_A`` test
        # (:127-129) only fires for foreign names; here an adapter without checkpoint weights is simply absent
        # (reset init has B == 0, i.e. it contributes exactly nothing), so the merged-default test comes first.
        if name == "default" and default_adapter_names is not None:
            subs = []
            for dn in default_adapter_names:
                xx = x.to(lora_A[dn].dtype)
                subs.append(F.linear(F.linear(xx, lora_A[dn]), lora_B[dn]) * scaling[dn])
            out["default"] = (base + torch.stack(subs).sum(0)).to(prev)
            continue
        if name not in lora_A:
            out[name] = base
            continue
        xx = x.to(lora_A[name].dtype)
        out[name] = (base + F.linear(F.linear(xx, lora_A[name]), lora_B[name]) * scaling[name]).to(prev)
    return out


def routed_sum(outputs: Dict[str, torch.Tensor], masks: Dict[str, torch.Tensor], like: torch.Tensor):
    """:266-268 — ``stack([out[k] * mask[k][..., None].to(x) for k in masks]).sum(0)``."""
    return torch.stack([outputs[k] * masks[k].unsqueeze(-1).to(like) for k in masks]).sum(dim=0)


class LinearParams:
    """One LocalLoraLinear's parameters: W [out,in], adapters name→(A [r,in], B [out,r])."""

    def __init__(self, W, lora_A, lora_B, scaling, default_adapter_names):
        self.W, self.lora_A, self.lora_B = W, lora_A, lora_B
        self.scaling, self.default_adapter_names = scaling, default_adapter_names

    def __call__(self, x, active_adapters):
        return lora_linear_forward(x, self.W, self.lora_A, self.lora_B, self.scaling, active_adapters,
                                   self.default_adapter_names)


# ----------------------------------------------------------------------------- norm / rope (transformers 4.31)
def rms_norm(x, weight, eps):
    """transformers LlamaRMSNorm: fp32 variance, rsqrt, cast back, × weight."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return weight * h.to(dt)


def rope_cos_sin(head_dim, seq_len, dtype, base=10000.0, cache_dtype=torch.float32, linear_factor=1.0):
    """transformers 4.31 LlamaRotaryEmbedding: inv_freq fp32, emb = cat(freqs, freqs); the cos/sin cache is
    stored in ``cache_dtype`` (``torch.get_default_dtype()`` at module construction) and cast to ``dtype`` on use.
    ``linear_factor``: LlamaLinearScalingRotaryEmbedding (config.rope_scaling = {"type": "linear", "factor": f},
    multimodal_llama.py:193-199) divides the positions by f before the outer product."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(seq_len, dtype=inv_freq.dtype)
    if linear_factor != 1.0:
        t = t / linear_factor
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(cache_dtype).to(dtype), emb.sin().to(cache_dtype).to(dtype)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin, position_ids):
    cos = cos[position_ids].unsqueeze(1)
    sin = sin[position_ids].unsqueeze(1)
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def causal_additive_mask(bsz, q_len, dtype, attention_mask_2d=None):
    """transformers 4.31 ``_prepare_decoder_attention_mask`` (multimodal_llama.py:543-545):
    additive finfo.min above the diagonal, plus expanded padding mask."""
    m = torch.full((q_len, q_len), torch.finfo(dtype).min, dtype=dtype).triu(1)
    m = m[None, None].expand(bsz, 1, q_len, q_len)
    if attention_mask_2d is not None:
        inv = 1.0 - attention_mask_2d[:, None, None, :].to(dtype).expand(bsz, 1, q_len, q_len)
        pad = inv.masked_fill(inv.to(torch.bool), torch.finfo(dtype).min)
        m = pad + m  # 4.31 adds the two masks
    return m


# ----------------------------------------------------------------------------- A14 attention, A15 MLP, A16 layer/model
def attention_forward(x, p: Dict[str, LinearParams], masks, modal_names, num_heads, position_ids, additive_mask,
                      rope_linear_factor=1.0):
    """:204-342 (no past_key_value, pretraining_tp == 1, num_key_value_heads == num_heads)."""
    bsz, q_len, hidden = x.shape
    hd = hidden // num_heads
    if masks is None:
        q = p["q_proj"](x, ("default",))["default"]
        k = p["k_proj"](x, ("default",))["default"]
        v = p["v_proj"](x, ("default",))["default"]
    else:
        q = routed_sum(p["q_proj"](x, modal_names), masks, x)
        k = routed_sum(p["k_proj"](x, modal_names), masks, x)
        v = routed_sum(p["v_proj"](x, modal_names), masks, x)
    q = q.view(bsz, q_len, num_heads, hd).transpose(1, 2)
    k = k.view(bsz, q_len, num_heads, hd).transpose(1, 2)
    v = v.view(bsz, q_len, num_heads, hd).transpose(1, 2)
    cos, sin = rope_cos_sin(hd, q_len, v.dtype, linear_factor=rope_linear_factor)
    q, k = apply_rope(q, k, cos, sin, position_ids)
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    if additive_mask is not None:
        w = w + additive_mask
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(bsz, q_len, hidden)
    if masks is None:
        return p["o_proj"](o, ("default",))["default"]
    return routed_sum(p["o_proj"](o, modal_names), masks, x)


def mlp_forward(x, p: Dict[str, LinearParams], masks, modal_names):
    """:380-394 — per adapter k: down_k(silu(gate_k(x)) * up_k(x)); masked sum."""
    if masks is None:
        g = F.silu(p["gate_proj"](x, ("default",))["default"])
        u = p["up_proj"](x, ("default",))["default"]
        return p["down_proj"](g * u, ("default",))["default"]
    gated = {k: F.silu(v) for k, v in p["gate_proj"](x, modal_names).items()}
    up = p["up_proj"](x, modal_names)
    down = {k: p["down_proj"](gated[k] * up[k], [k])[k] for k in up}
    return routed_sum(down, masks, x)


def decoder_layer_forward(x, layer, masks, modal_names, num_heads, position_ids, additive_mask, eps, rope_linear_factor=1.0):
    """:408-468.  ``layer``: dict with 'input_layernorm', 'post_attention_layernorm' (weights) and the 7 LinearParams."""
    h = rms_norm(x, layer["input_layernorm"], eps)
    h = attention_forward(h, layer, masks, modal_names, num_heads, position_ids, additive_mask, rope_linear_factor)
    x = x + h
    h = rms_norm(x, layer["post_attention_layernorm"], eps)
    h = mlp_forward(h, layer, masks, modal_names)
    return x + h


def model_forward(inputs_embeds, layers, final_norm, lm_head, masks, modal_names, num_heads, eps,
                  attention_mask_2d=None, rope_linear_factor=1.0):
    """:526-545 (position ids, mask), :561-603 (layer loop, final norm), :720 (lm_head on all positions)."""
    bsz, q_len, _ = inputs_embeds.shape
    position_ids = torch.arange(q_len, dtype=torch.long).unsqueeze(0)
    if attention_mask_2d is None:
        attention_mask_2d = torch.ones((bsz, q_len), dtype=torch.bool)
    additive = causal_additive_mask(bsz, q_len, inputs_embeds.dtype, attention_mask_2d)
    h = inputs_embeds
    for layer in layers:
        h = decoder_layer_forward(h, layer, masks, modal_names, num_heads, position_ids, additive, eps, rope_linear_factor)
    h = rms_norm(h, final_norm, eps)
    return F.linear(h, lm_head), h
