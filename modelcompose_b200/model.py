"""Composed-model prefill: host mirror of the reference model classes over the CUDA library.

Mirrors (all paths relative to the reference root):
  modelcompose/model/language_model/multimodal_llama.py   MultimodalConfig (:33-61), LocalLoraLinear coefficient
      handling (:70-118), routed attention / MLP (:262-268, :335-336, :380-390), decoder layer / model / CausalLM
      forward (:408-468, :488-619, :676-745)
  modelcompose/model/multimodal_encoder/builder.py:119-130   infer_modals
  modelcompose/model/multimodal_arch.py:197-268              encode_modal_inputs (projector + prefix/suffix)
  modelcompose/model/multimodal_projector/builder.py:202-219 projector types

Every matrix product, the causal attention of the prefill, the splice, RMSNorm, RoPE and SiLU·mul run in
``libmodelcompose_b200.so`` (no torch fallback; a missing library raises).  The modality
encoders are frozen third-party feature extractors outside the hot path: ``modal_inputs`` here
carries their OUTPUT features (``[b, n, d]``, video ``[b, t, n, d]``), i.e. what ``encoder(inputs)`` returns at
multimodal_arch.py:231-243.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _cabi
from . import decode as DC
from . import linear as LN
from . import materialize as MZ
from . import splice as SP

ADAPTER_ORDER = ("audio", "vision", "video", "point")
# kernel variant of the base + LoRA-up launches (mc_linear_plan_create `tuning`): 0 = 128x256 single-CTA tiles, 3 = 512x256
# CTA-pair tiles (cta_group::2), 4 = 256x256 CTA-pair tiles; all bit-identical (tests/test_prefill_gpu.py).  Unset = auto: the
# 512x256 pair kernel once the batch has enough rows to fill 74 CTA pairs several times over — it takes the same time for the
# linears at lower power, and the power-capped step as a whole runs 2 % faster (profiles/r01_linear_pair_instep.txt) — and the
# single-CTA kernel below that.
_UP_TUNING_ENV = os.environ.get("MC_LINEAR_UP_TUNING")
UP_TUNING = int(_UP_TUNING_ENV) if _UP_TUNING_ENV not in (None, "", "auto") else None
UP_TUNING_PAIR_MIN_ROWS = 8192
# development switches: tile rasterisation of the base ("up") launches, M tiles per sweep over N (0 = the library's choice), for the
# CTA-pair kernel (512-row tiles) and the single-CTA kernel (128-row tiles); swept in profiles/r02_raster.txt
GROUP_M_PAIR = int(os.environ.get("MC_LINEAR_GROUP_M_PAIR", "0"))
GROUP_M_SINGLE = int(os.environ.get("MC_LINEAR_GROUP_M_SINGLE", "0"))
# ... except for the launches whose epilogue is heavy or whose K is short: the pair kernel's accumulator is single-buffered (its
# epilogue is not overlapped with the next tile), and ncu shows its tensor pipe at 63 % on the up_proj launch that carries
# SiLU(gate)·up and 67 % on o_proj, against 78 - 81 % on gate_proj / down_proj (profiles/r01_linear_pair_ncu_instep.txt).
# Those two stay on the single-CTA kernel, whose epilogue overlaps the next tile's MMAs.
UP_AUTO_SINGLE_CTA = ("up_u", "up_o")
FUSE_ROPE = os.environ.get("MC_FUSE_ROPE", "1") != "0"  # development switch: 0 = separate mc_rope launch
# Prefill activations live in MODALITY-MAJOR row order (all text rows of the batch, then all audio rows, ...), so that every
# 128-row tile of the routed linears holds one adapter group; only attention sees sequence order (the q / k / v epilogues
# scatter rows back, the attention output is gathered again).  Development switch: 0 = sequence order everywhere.
MODALITY_MAJOR = os.environ.get("MC_MODALITY_MAJOR", "1") != "0"
# causal prefill attention: 1 (default) = this library's tcgen05 kernel (head_dim 128); 0 = the stock cuDNN / flash-attn call,
# kept as a development switch for A/B timing only
ATTENTION_NATIVE = os.environ.get("MC_ATTENTION_NATIVE", "1") != "0"
# Evaluation form of the routed linears.  0 (default): the reference's own form — base weight plus the low-rank branch of the
# token's group, evaluated per forward (LocalLoraLinear.forward, multimodal_llama.py:130-149).  1: MATERIALISED — at load one
# dense W_eff,g per routing group is built on the device (materialize.py: rank-r GEMMs + the N-source merge kernel for the
# online-merge-reset blend of the text group) and every linear becomes one grouped GEMM over the modality-major rows; costs
# one extra copy of the decoder weights per group in HBM, removes the four LoRA-down launches per layer and the rank-space
# traffic.  W_eff is rounded once to the storage dtype, so logits differ from form 0 within the bar stated in
# tests/test_prefill_gpu.py (reference tooling for the dense form: delta_weights_compare.py:24-31,61).
MATERIALIZE = os.environ.get("MC_MATERIALIZE", "0") != "0"
# Branch form for the prefill, dense weights for the DECODE steps only: generated tokens are text, so every decode row takes the
# default group and its W_eff (W + the default adapters) is all the decode step needs — one extra copy of the decoder weights
# (13 GB for vicuna-7B) instead of one per group, four launches per layer fewer in the HBM-bound step (+40 % tokens/s at
# batch 32, profiles/r02_decode.txt).  Decode logits then carry the dense form's single rounding of W_eff, like MC_MATERIALIZE=1.
DECODE_DENSE = os.environ.get("MC_DECODE_DENSE", "0") != "0"


class MultimodalConfig:
    """Attribute bag with the reference defaults (multimodal_llama.py:33-61) over LLaMA's config fields."""
    model_type = "multimodal"
    lora_strategy = None
    lora_name = "default"
    lora_r = 128
    lora_alpha = 256
    lora_dropout = 0.05
    local_prefix_tokens = 0
    local_suffix_tokens = 0
    merge_default_weights = None
    reset_scaling_weights = None
    mm_vision_encoder = None
    mm_vision_tower = None
    mm_audio_encoder = None
    mm_video_encoder = None
    mm_point_encoder = None
    hidden_size = 4096
    intermediate_size = 11008
    num_attention_heads = 32
    num_key_value_heads = None
    num_hidden_layers = 32
    vocab_size = 32000
    rms_norm_eps = 1e-6
    max_position_embeddings = 2048
    rope_theta = 10000.0
    hidden_act = "silu"

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads

    @classmethod
    def from_dict(cls, d: dict) -> "MultimodalConfig":
        return cls(**d)

    def to_dict(self) -> dict:
        return {k: v for k, v in vars(self).items()}


def infer_modals(config) -> List[str]:
    """multimodal_encoder/builder.py:119-130 — adapter / modality order (fixes the summation order of A9)."""
    modals = ["default"]
    if getattr(config, "mm_audio_encoder", None) is not None:
        modals.append("audio")
    if getattr(config, "mm_vision_encoder", None) is not None or getattr(config, "mm_vision_tower", None) is not None:
        modals.append("vision")
    if getattr(config, "mm_video_encoder", None):
        modals.append("video")
    if getattr(config, "mm_point_encoder", None):
        modals.append("point")
    return modals


def extract_params(input_string: str) -> Dict[str, float]:
    """multimodal_llama.py:109-118."""
    params = {}
    for pair in input_string.split(","):
        key, value = pair.split("=")
        params[key.strip()] = float(value)
    return params


def adapter_scaling(modal_names: Sequence[str], r: int, lora_alpha: float, reset_scaling_weights: Optional[str]):
    """multimodal_llama.py:84-106 — (adapter names, scaling dict, default_adapter_names or None); float64 like Python."""
    names = list(modal_names)
    scaling = {n: lora_alpha / r for n in names}
    default_adapter_names = None
    if reset_scaling_weights is not None:
        reset = extract_params(reset_scaling_weights)
        if any("default-" in k for k in reset):
            default_adapter_names = [f"default-{n}" for n in names[1:]]
            for n in default_adapter_names:
                names.append(n)
                scaling[n] = lora_alpha / r
        for k in reset:
            if k in scaling:
                scaling[k] = scaling[k] * reset[k]
    return names, scaling, default_adapter_names


@dataclass
class CausalLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[list] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None
    modal_id: Optional[torch.Tensor] = None  # extra: uint8 [B, S'] routing ids (0 = default)


DECODE_GRAPH = os.environ.get("MC_DECODE_GRAPH", "1") != "0"    # 0: launch the decode step's kernels one by one (development)
DECODE_NATIVE = os.environ.get("MC_DECODE_NATIVE", "1") != "0"  # 0: decode through the prefill kernels at M = batch (development)

LINEARS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def _linear_key(layer: int, name: str) -> str:
    block = "self_attn" if name in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
    return f"model.layers.{layer}.{block}.{name}"


@dataclass
class _Layer:
    W: Dict[str, torch.Tensor] = field(default_factory=dict)
    ad: Dict[str, LN.PackedAdapters] = field(default_factory=dict)
    Weff: Dict[str, List[torch.Tensor]] = field(default_factory=dict)  # materialised form: one dense weight per routing group
    Wdec: Dict[str, torch.Tensor] = field(default_factory=dict)        # decode_dense: W_eff of the text group, for the decode step only
    ln1: torch.Tensor = None
    ln2: torch.Tensor = None


class KVCache:
    """Key/value cache of one batch: per layer ``[B, heads, capacity, head_dim]`` (keys stored after RoPE, as transformers 4.31
    caches them, multimodal_llama.py:284-289) — the reference's ``past_key_value`` layout with room to grow, so the keys of
    one (sequence, head) are one contiguous stream for the decode attention.  Returned as ``past_key_values`` when ``use_cache``."""

    def __init__(self, n_layers: int, B: int, capacity: int, n_heads: int, head_dim: int, dtype, device):
        self.k = [torch.empty((B, n_heads, capacity, head_dim), dtype=dtype, device=device) for _ in range(n_layers)]
        self.v = [torch.empty((B, n_heads, capacity, head_dim), dtype=dtype, device=device) for _ in range(n_layers)]
        self.length = 0
        self.capacity = capacity
        self.prefill_mask = None

    def grow(self, capacity: int) -> None:
        for lst in (self.k, self.v):
            for i, t in enumerate(lst):
                n = torch.empty((t.shape[0], t.shape[1], capacity, t.shape[3]), dtype=t.dtype, device=t.device)
                n[:, :, :self.length].copy_(t[:, :, :self.length])
                lst[i] = n
        self.capacity = capacity

    def legacy(self):
        """transformers-4.31 layout: tuple over layers of (k, v) ``[B, heads, length, head_dim]`` views."""
        return tuple((k[:, :, :self.length], v[:, :, :self.length]) for k, v in zip(self.k, self.v))

    def __len__(self):
        return len(self.k)

    def __getitem__(self, i):
        return self.legacy()[i]


class _Workspace:
    """Static activation buffers and launch plans for one (batch, padded length) shape."""

    def __init__(self, model: "MultimodalLlamaForCausalLM", B: int, S: int):
        cfg, dev, dt = model.config, model.device, model.dtype
        T, H, I, V = B * S, cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
        R = model.rank_total
        self.B, self.S, self.T = B, S, T
        # a decode step has at most 128 rows: 128x128 tiles double the number of CTAs streaming the weights
        self.up_tuning = 1 if T <= LN.TILE_M else (UP_TUNING if UP_TUNING is not None else (3 if T >= UP_TUNING_PAIR_MIN_ROWS else 0))
        self.up_mixed = UP_TUNING is None and self.up_tuning == 3  # auto: per-launch choice (UP_AUTO_SINGLE_CTA)

        def buf(*shape, dtype=dt):
            return torch.empty(shape, dtype=dtype, device=dev)
        self.x = buf(T, H)
        self.xn = buf(T, H)
        self.q, self.k, self.v, self.attn = buf(T, H), buf(T, H), buf(T, H), buf(T, H)
        self.dense = model.materialize
        self.t = [] if self.dense else [buf(T, R) for _ in range(3)]
        # materialised form: first buffer row of every routing group (rows are modality-major), rewritten per batch
        self.seg_start = torch.zeros(len(model.modal_names) + 1, dtype=torch.int32, device=dev)
        self.seg_start[1:] = T
        self.gate = buf(T, I)
        self.logits = buf(T, V)
        self.row_group = torch.zeros(T, dtype=torch.uint8, device=dev)
        # perm[i] = sequence-order row held at row i of the activation buffers (identity for the decode step / unrouted)
        self.permute = MODALITY_MAJOR and S > 1
        self.perm = torch.arange(T, dtype=torch.int32, device=dev) if self.permute else None
        self.inv_perm = torch.arange(T, dtype=torch.int32, device=dev) if self.permute else None  # buffer row of sequence row t
        self.perm_is_identity = True
        # RoPE rides in the q / k projection epilogue when a tile holds whole heads; otherwise mc_rope runs after it
        D = H // cfg.num_attention_heads
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pos_value = 0
        self.rope_fused = FUSE_ROPE and D in (64, 128, 256) and (128 if T <= LN.TILE_M or H < 256 else 256) % D == 0
        self.rope = (model._rope[0], model._rope[1], self.pos, S, D) if self.rope_fused else None
        self.mtile = torch.zeros((T + LN.TILE_M - 1) // LN.TILE_M, dtype=torch.int32, device=dev)
        self.plans: List[Dict[str, LN.LinearPlan]] = []
        if self.dense and MODALITY_MAJOR is False and S > 1:
            raise ValueError("the materialised form needs the modality-major row order (MC_MODALITY_MAJOR=1)")
        for layer in model.layers:
            self.plans.append(self._layer_plans_dense(layer) if self.dense else self._layer_plans(layer))
        self.lm_head = LN.LinearPlan([LN.Problem(self.xn, model.lm_head, self.logits, c_rowmap=self.perm)], tuning=self.up_tuning)
        # generation only consumes the last position's logits (the reference computes all S' and slices, :720 + HF generate):
        # a B-row lm_head over the gathered last rows
        self.last_idx = torch.arange(B, dtype=torch.int32, device=dev) * S + (S - 1)
        self.xn_last = buf(B, H)
        self.logits_last = buf(B, V)
        self.lm_head_last = LN.LinearPlan([LN.Problem(self.xn_last, model.lm_head, self.logits_last)], tuning=1)

    def set_routing(self, modal_id: Optional[torch.Tensor], lut: Optional[Sequence[int]] = None) -> Optional[torch.Tensor]:
        """Routing of this batch: ``modal_id`` ``[B, S]`` uint8 in sequence order (None = all default) with ``lut`` mapping its
        values to routing groups (None = they are routing groups already) -> row permutation, per-row groups in buffer order,
        segment table / per-tile group masks.  Returns the routing groups in sequence order (``[B, S]`` uint8) or None."""
        n_groups = self.seg_start.numel() - 1
        if modal_id is None:
            self.row_group.zero_()
            if self.permute and not self.perm_is_identity:
                self.perm.copy_(torch.arange(self.T, dtype=torch.int32, device=self.perm.device))
                self.inv_perm.copy_(self.perm)
                self.perm_is_identity = True
            self.seg_start[1:] = self.T
            group_seq = None
        elif self.permute:
            # stable counting sort by group on the device (mc_route_permutation): one launch, no host synchronisation
            group_seq = torch.empty((self.B, self.S), dtype=torch.uint8, device=self.perm.device)
            LN.route_permutation(modal_id.reshape(-1).contiguous(), lut, n_groups, self.perm, self.inv_perm, self.row_group, self.seg_start,
                                 group_seq.view(-1))
            self.perm_is_identity = False
        else:
            if self.dense:
                raise ValueError("the materialised form needs rows grouped by modality: route in modality-major order")
            group_seq = modal_id if lut is None else torch.tensor(list(lut), dtype=torch.uint8, device=modal_id.device)[modal_id.long()]
            self.row_group.view(self.B, self.S).copy_(group_seq)
        if not self.dense:
            LN.route_tile_masks(self.row_group, self.mtile, coarsen={3: 4, 4: 2}.get(self.up_tuning & 0xff, 1))
        return group_seq

    def last_rows(self) -> torch.Tensor:
        """int32 [B]: buffer row holding the last position of every sequence."""
        if self.permute and not self.perm_is_identity:
            return self.inv_perm[self.last_idx.long()].contiguous()
        return self.last_idx

    def load_rows(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        """dst (buffer order) <- src (sequence order), both [T, width]."""
        if src.dtype != dst.dtype:
            src = src.to(dst.dtype)
        if self.permute and not self.perm_is_identity:
            LN.gather_rows(src, self.perm, dst)
        else:
            dst.copy_(src)

    def sequence_order(self, buf: torch.Tensor) -> torch.Tensor:
        """A copy of an activation buffer in sequence order ``[B, S, width]`` (hidden-state outputs, tests)."""
        out = torch.empty_like(buf)
        if self.permute and not self.perm_is_identity:
            LN.gather_rows(buf, self.inv_perm, out)
        else:
            out.copy_(buf)
        return out.view(self.B, self.S, -1)

    def _down(self, src, layer: _Layer, names, tbufs):
        return LN.LinearPlan([LN.Problem(src, layer.ad[n].A_all, t, col_scale=layer.ad[n].col_scale, row_group=self.row_group,
                                         mtile_mask=self.mtile, group_cols=layer.ad[n].group_cols, epilogue=LN.EPI_ROWMASK)
                              for n, t in zip(names, tbufs)], tuning=1)  # 128x128 tiles: one routing group per N tile

    def _up(self, src, layer: _Layer, names, tbufs, outs, residual=None, epilogue=None, launch=None):
        tuning = 0 if (self.up_mixed and launch in UP_AUTO_SINGLE_CTA) else self.up_tuning
        tuning |= ((GROUP_M_PAIR if tuning in (3, 4) else GROUP_M_SINGLE) & 0xff) << 8
        if epilogue is None:
            epilogue = LN.EPI_RESIDUAL if residual is not None else LN.EPI_NONE
        probs = []
        for n, t, o in zip(names, tbufs, outs):
            rope = self.rope if n in ("q_proj", "k_proj") else None
            # q / k / v leave the modality-major order here: attention needs [B, S, heads, D]
            rowmap = self.perm if n in ("q_proj", "k_proj", "v_proj") else None
            probs.append(LN.Problem(src, layer.W[n], o, A1=t, B1=layer.ad[n].B_all, mtile_mask=self.mtile,
                                    group_cols=layer.ad[n].group_cols, residual=residual,
                                    epilogue=LN.EPI_ROPE if rope is not None else epilogue, rope=rope, c_rowmap=rowmap))
        return LN.LinearPlan(probs, tuning=tuning)

    def _dense(self, src, layer: _Layer, names, outs, residual=None, epilogue=None, launch=None):
        """Grouped GEMM over the materialised weights: rows of routing group g (a contiguous segment of the modality-major
        buffers) multiply with W_eff,g."""
        tuning = 0 if (self.up_mixed and launch in UP_AUTO_SINGLE_CTA) else self.up_tuning
        if tuning == 4:
            tuning = 3
        if epilogue is None:
            epilogue = LN.EPI_RESIDUAL if residual is not None else LN.EPI_NONE
        probs = []
        for n, o in zip(names, outs):
            rope = self.rope if n in ("q_proj", "k_proj") else None
            rowmap = self.perm if n in ("q_proj", "k_proj", "v_proj") else None
            probs.append(LN.Problem(src, layer.W[n], o, residual=residual, epilogue=LN.EPI_ROPE if rope is not None else epilogue,
                                    rope=rope, c_rowmap=rowmap, seg_start=self.seg_start, B0_groups=layer.Weff[n]))
        return LN.LinearPlan(probs, tuning=tuning)

    def _layer_plans_dense(self, layer: _Layer) -> Dict[str, LN.LinearPlan]:
        qkv = ("q_proj", "k_proj", "v_proj")
        return {
            "up_qkv": self._dense(self.xn, layer, qkv, (self.q, self.k, self.v)),
            "up_o": self._dense(self.attn, layer, ("o_proj",), (self.x,), residual=self.x, launch="up_o"),
            "up_g": self._dense(self.xn, layer, ("gate_proj",), (self.gate,)),
            "up_u": self._dense(self.xn, layer, ("up_proj",), (self.gate,), residual=self.gate, epilogue=LN.EPI_SILU_MUL, launch="up_u"),
            "up_d": self._dense(self.gate, layer, ("down_proj",), (self.x,), residual=self.x),
        }

    def _layer_plans(self, layer: _Layer) -> Dict[str, LN.LinearPlan]:
        qkv, gu = ("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj")
        return {
            "down_qkv": self._down(self.xn, layer, qkv, self.t),
            "up_qkv": self._up(self.xn, layer, qkv, self.t, (self.q, self.k, self.v)),
            "down_o": self._down(self.attn, layer, ("o_proj",), self.t[:1]),
            "up_o": self._up(self.attn, layer, ("o_proj",), self.t[:1], (self.x,), residual=self.x, launch="up_o"),
            "down_gu": self._down(self.xn, layer, gu, self.t[:2]),
            # gate first, then up with the SiLU·mul folded into its epilogue (act overwrites the gate buffer)
            "up_g": self._up(self.xn, layer, gu[:1], self.t[:1], (self.gate,)),
            "up_u": self._up(self.xn, layer, gu[1:], self.t[1:2], (self.gate,), residual=self.gate, epilogue=LN.EPI_SILU_MUL,
                             launch="up_u"),
            "down_d": self._down(self.gate, layer, ("down_proj",), self.t[:1]),
            "up_d": self._up(self.gate, layer, ("down_proj",), self.t[:1], (self.x,), residual=self.x),
        }


class MultimodalLlamaForCausalLM:
    """Inference-only drop-in for the reference class of the same name (multimodal_llama.py:622-767)."""

    def __init__(self, config: MultimodalConfig, base_state_dict: Dict[str, torch.Tensor],
                 adapter_state_dict: Optional[Dict[str, torch.Tensor]] = None, device="cuda", dtype=torch.float16,
                 materialize: Optional[bool] = None, decode_dense: Optional[bool] = None):
        _cabi.lib()  # fail loudly if the CUDA library is missing
        self.materialize = MATERIALIZE if materialize is None else bool(materialize)
        self.decode_dense = (DECODE_DENSE if decode_dense is None else bool(decode_dense)) and not self.materialize
        if dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("inference dtype must be float16 (reference builder.py:185) or bfloat16")
        self.config, self.device, self.dtype = config, torch.device(device), dtype
        # config.rope_scaling (multimodal_llama.py:190-203): "linear" divides the positions by the factor — a different table, nothing
        # else changes.  "dynamic" (NTK) re-derives the base from the longest sequence seen so far, per call and per decode step,
        # also for keys already in the cache: a table per step does not fit the captured decode step, so it is refused.
        self.rope_linear_factor = 1.0
        scaling_cfg = getattr(config, "rope_scaling", None)
        if scaling_cfg is not None:
            kind, factor = scaling_cfg.get("type"), float(scaling_cfg.get("factor", 1.0))
            if kind == "linear":
                if not factor >= 1.0:
                    raise ValueError(f"rope_scaling factor must be >= 1, got {factor}")
                self.rope_linear_factor = factor
            elif kind == "dynamic":
                raise NotImplementedError("rope_scaling type 'dynamic' (NTK, multimodal_llama.py:200-203) is not implemented on this path; "
                                          "'linear' is")
            else:
                raise ValueError(f"Unknown RoPE scaling type {kind}")  # the reference's message (multimodal_llama.py:205)
        if int(getattr(config, "pretraining_tp", 1) or 1) > 1:
            raise NotImplementedError("pretraining_tp > 1 (sliced projections that bypass the adapters, multimodal_llama.py:222-237, "
                                      ":323-326, :366-377) is not implemented on this path")
        if config.num_key_value_heads != config.num_attention_heads:
            raise NotImplementedError(f"grouped-query attention (num_key_value_heads {config.num_key_value_heads} != "
                                      f"num_attention_heads {config.num_attention_heads}) is not implemented on this path")
        self.modal_names = infer_modals(config)
        sd = adapter_state_dict or {}
        r, alpha = config.lora_r, config.lora_alpha
        self.adapter_names, self.scaling, self.default_adapter_names = adapter_scaling(
            self.modal_names, r, alpha, config.reset_scaling_weights)

        def dev(t):
            return t.to(device=self.device, dtype=dtype).contiguous()
        H = config.hidden_size
        self.embed_tokens = dev(base_state_dict["model.embed_tokens.weight"])
        self.norm = dev(base_state_dict["model.norm.weight"])
        self.lm_head = dev(base_state_dict["lm_head.weight"])
        self.layers: List[_Layer] = []
        for li in range(config.num_hidden_layers):
            layer = _Layer()
            layer.ln1 = dev(base_state_dict[f"model.layers.{li}.input_layernorm.weight"])
            layer.ln2 = dev(base_state_dict[f"model.layers.{li}.post_attention_layernorm.weight"])
            for name in LINEARS:
                key = _linear_key(li, name)
                W = dev(base_state_dict[key + ".weight"])
                layer.W[name] = W
                # adapters present in the checkpoint; absent ones keep the loader's reset init (B = 0, builder.py:150-153)
                # and contribute exactly nothing, so they are simply left out of the packed layout
                A = {a: dev(sd[f"{key}.lora_A.{a}.weight"]) for a in self.adapter_names if f"{key}.lora_A.{a}.weight" in sd}
                Bm = {a: dev(sd[f"{key}.lora_B.{a}.weight"]) for a in A}
                if self.materialize:
                    layer.Weff[name] = MZ.effective_weights(W, A, Bm, self.scaling, self.modal_names, self.default_adapter_names,
                                                            alpha / r)
                else:
                    layer.ad[name] = LN.pack_adapters(A, Bm, self.scaling, self.modal_names, self.default_adapter_names,
                                                      W.shape[1], W.shape[0], dtype, self.device)
                    if self.decode_dense:  # W_eff of the text group only
                        layer.Wdec[name] = MZ.effective_weights(W, A, Bm, self.scaling, self.modal_names[:1], self.default_adapter_names,
                                                                alpha / r)[0]
            self.layers.append(layer)
        self.rank_total = self.layers[0].ad["q_proj"].A_all.shape[0] if self.layers and not self.materialize else LN.K_BLOCK
        for layer in self.layers:
            for name in LINEARS:
                if not self.materialize and layer.ad[name].A_all.shape[0] != self.rank_total:
                    raise ValueError("every linear must carry the same adapter ranks")

        # projectors (multimodal_projector/builder.py:202-219): 'linear' or 'mlp{N}x_gelu'
        self.projectors: Dict[str, List[Tuple[torch.Tensor, torch.Tensor]]] = {}
        for modal in self.modal_names[1:]:
            stem = f"model.modal_projectors.{modal}"
            if f"{stem}.weight" in sd:
                self.projectors[modal] = [(dev(sd[f"{stem}.weight"]), dev(sd[f"{stem}.bias"]))]
            else:
                layers, i = [], 0
                while f"{stem}.{i}.weight" in sd:
                    layers.append((dev(sd[f"{stem}.{i}.weight"]), dev(sd[f"{stem}.{i}.bias"])))
                    i += 2
                if layers:
                    self.projectors[modal] = layers
        # prefix / suffix tokens (multimodal_llama.py:634-649): zeros unless the checkpoint carries them
        self.prefix_tokens = self._local_tokens(sd, "prefix", config.local_prefix_tokens)
        self.suffix_tokens = self._local_tokens(sd, "suffix", config.local_suffix_tokens)
        self._attn_backend: Optional[str] = os.environ.get("MC_ATTENTION_BACKEND") or None
        self._ws: Dict[Tuple[int, int], _Workspace] = {}
        self._rope: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._proj_cache: Dict[tuple, tuple] = {}
        self._dws: Optional["DC.DecodeWorkspace"] = None

    # ------------------------------------------------------------------------------------------ construction helpers
    def _local_tokens(self, sd, kind: str, n_default: int):
        if not n_default:
            return None
        out = {}
        for modal in self.modal_names:
            n = getattr(self.config, f"local_{modal}_{kind}_tokens", None)
            n = n_default if n is None else n
            key = f"{kind}_tokens.{modal}"
            if key in sd:
                out[modal] = sd[key].to(device=self.device, dtype=self.dtype).contiguous()
            else:
                out[modal] = torch.zeros((1, n, self.config.hidden_size), dtype=self.dtype, device=self.device)
        return out

    def _rope_tables(self, seq_len: int):
        """transformers 4.31 LlamaRotaryEmbedding / LlamaLinearScalingRotaryEmbedding: fp32 cache built on the host, cast to the model dtype on use.  The
        tables are referenced by the launch plans, so they only ever grow (in 4096-position steps) and growing drops the
        cached workspaces."""
        n = max(seq_len, self.config.max_position_embeddings)
        if self._rope is None or self._rope[0].shape[0] < n:
            n = (n + 4095) // 4096 * 4096
            D = self.config.hidden_size // self.config.num_attention_heads
            base = float(getattr(self.config, "rope_theta", 10000.0) or 10000.0)
            inv_freq = 1.0 / (base ** (torch.arange(0, D, 2).float() / D))
            t = torch.arange(n, dtype=inv_freq.dtype)
            if self.rope_linear_factor != 1.0:
                t = t / self.rope_linear_factor  # LlamaLinearScalingRotaryEmbedding
            freqs = torch.einsum("i,j->ij", t, inv_freq)
            emb = torch.cat((freqs, freqs), dim=-1)
            self._rope = (emb.cos().to(self.dtype).to(self.device).contiguous(),
                          emb.sin().to(self.dtype).to(self.device).contiguous())
            self._ws.clear()
        return self._rope

    def get_model(self):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------------------------------ encode_modal_inputs
    def project_modal_features(self, modal_inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """multimodal_arch.py:197-243 after the (frozen, out-of-scope) encoders: one grouped launch per projector depth."""
        feats, order = {}, [m for m in self.modal_names[1:] if m in modal_inputs]
        for m in order:
            f = modal_inputs[m]
            if f.dim() == 4:  # video [b, t, n, d] -> [b, t*n, d] (:236-240)
                f = f.reshape(f.shape[0], f.shape[1] * f.shape[2], f.shape[3])
            if f.dim() != 3:
                raise ValueError(f"modal_inputs[{m}] must be encoder features [b, n, d] (video [b, t, n, d])")
            if m not in self.projectors:
                raise KeyError(f"no projector weights for modality {m}")
            feats[m] = f.to(device=self.device, dtype=self.dtype).contiguous()
        if not order:
            return {}
        key = tuple((m, tuple(feats[m].shape), feats[m].data_ptr()) for m in order)
        if key not in self._proj_cache:
            self._proj_cache.clear()
            depth = max(len(self.projectors[m]) for m in order)
            cur = {m: feats[m].view(-1, feats[m].shape[-1]) for m in order}
            plans, keep = [], []
            for d in range(depth):
                probs, nxt = [], dict(cur)
                for m in order:
                    if d >= len(self.projectors[m]):
                        continue
                    W, b = self.projectors[m][d]
                    out = torch.empty((cur[m].shape[0], W.shape[0]), dtype=self.dtype, device=self.device)
                    last = d == len(self.projectors[m]) - 1
                    probs.append(LN.Problem(cur[m], W, out, bias=b, epilogue=LN.EPI_BIAS if last else LN.EPI_BIAS_GELU))
                    nxt[m] = out
                    keep.append(out)
                for i in range(0, len(probs), LN.MAX_PROBLEMS):
                    plans.append(LN.LinearPlan(probs[i:i + LN.MAX_PROBLEMS]))
                cur = nxt
            outs = {m: cur[m].view(feats[m].shape[0], feats[m].shape[1], -1) for m in order}
            self._proj_cache[key] = (plans, outs, [feats[m] for m in order])
        plans, outs, _ = self._proj_cache[key]
        for p in plans:
            p.run()
        return outs

    # ------------------------------------------------------------------------------------------ forward
    def _rmsnorm(self, x, w, out):
        _cabi.check(_cabi.lib().mc_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0),
                                           out.stride(0), float(self.config.rms_norm_eps), _cabi.dtype_code(self.dtype),
                                           _cabi.current_stream_ptr()), "mc_rmsnorm")
        _cabi.count_launch()

    def _causal_attention(self, q, k, v, scale):
        """Stock-library causal attention on [B, S, heads, D].  Preference order measured on B200
        (tools/bench_attention_lib.py): cuDNN fused attention through torch SDPA (0.36 ms at C3 shapes), flash-attn 2
        (0.86 ms), torch's default SDPA choice.  The first backend that works is remembered."""
        F = torch.nn.functional
        order = [self._attn_backend] if self._attn_backend else ["cudnn", "flash_attn", "sdpa"]
        last_err = None
        for name in order:
            try:
                if name == "cudnn":
                    from torch.nn.attention import SDPBackend, sdpa_kernel
                    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
                        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                           is_causal=True, scale=scale).transpose(1, 2)
                elif name == "flash_attn":
                    from flash_attn import flash_attn_func
                    o = flash_attn_func(q, k, v, causal=True, softmax_scale=scale)
                else:
                    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                       is_causal=True, scale=scale).transpose(1, 2)
                self._attn_backend = name
                return o
            except (ImportError, RuntimeError) as e:
                last_err = e
        raise RuntimeError(f"no causal attention backend available: {last_err}")

    def _attention(self, ws: _Workspace, attention_mask: Optional[torch.Tensor], full: bool,
                   cache: Optional[KVCache] = None, layer_idx: int = 0, past: int = 0):
        """RoPE + causal attention of one layer.  ``past`` > 0 is the decode step: the new keys/values are appended to
        ``cache`` and the queries attend to everything cached (multimodal_llama.py:274-312 with past_key_value)."""
        cfg = self.config
        nH = cfg.num_attention_heads
        D = cfg.hidden_size // nH
        B, S = ws.B, ws.S
        if not ws.rope_fused:
            cos, sin = self._rope
            _cabi.check(_cabi.lib().mc_rope(ws.q.data_ptr(), ws.k.data_ptr(), cos.data_ptr(), sin.data_ptr(), ws.T, S, past, nH, D,
                                            ws.q.stride(0), ws.k.stride(0), _cabi.dtype_code(self.dtype),
                                            _cabi.current_stream_ptr()), "mc_rope")
            _cabi.count_launch()
        q, k, v = (t.view(B, S, nH, D) for t in (ws.q, ws.k, ws.v))
        if cache is not None:
            cache.k[layer_idx][:, :, past:past + S].copy_(k.transpose(1, 2))
            cache.v[layer_idx][:, :, past:past + S].copy_(v.transpose(1, 2))
        F = torch.nn.functional
        if past > 0:
            kk, vv = cache.k[layer_idx][:, :, :past + S], cache.v[layer_idx][:, :, :past + S]  # [B, heads, L, D]
            mask = None
            if not full or S > 1:
                neg = torch.finfo(self.dtype).min
                mask = torch.zeros((B, 1, S, past + S), dtype=self.dtype, device=self.device)
                if S > 1:
                    mask = mask + torch.full((S, past + S), neg, dtype=self.dtype, device=self.device).triu(past + 1)[None, None]
                if not full:
                    mask = (mask + (~attention_mask.bool())[:, None, None, :].to(self.dtype) * neg).clamp_min(neg)
            o = F.scaled_dot_product_attention(q.transpose(1, 2), kk, vv, attn_mask=mask, scale=1.0 / math.sqrt(D)).transpose(1, 2)
        elif full and D == 128 and ATTENTION_NATIVE:
            # own tcgen05 kernel: reads the projection outputs in place and writes straight into buffer (modality-major) order
            LN.attention_causal(ws.q, ws.k, ws.v, ws.attn, B, S, nH, 1.0 / math.sqrt(D),
                                out_rowmap=ws.inv_perm if ws.permute and not ws.perm_is_identity else None)
            return
        elif full:
            o = self._causal_attention(q, k, v, 1.0 / math.sqrt(D))
        else:
            # padded batch: additive mask as transformers 4.31 _prepare_decoder_attention_mask builds it (:543-545)
            neg = torch.finfo(self.dtype).min
            causal = torch.full((S, S), neg, dtype=self.dtype, device=self.device).triu(1)[None, None]
            pad = (~attention_mask.bool())[:, None, None, :].to(self.dtype) * neg
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                               attn_mask=(causal + pad).clamp_min(neg)).transpose(1, 2)
        ws.load_rows(o.reshape(B * S, nH * D), ws.attn)  # back to the buffers' row order for o_proj

    def prefill(self, inputs_embeds: torch.Tensor, modal_id: Optional[torch.Tensor], attention_mask=None,
                use_cache: bool = False, output_hidden_states: bool = False, past_key_values: Optional[KVCache] = None,
                last_logits_only: bool = False, cache_extra: int = 128, modal_lut: Optional[Sequence[int]] = None):
        """MultimodalLlamaModel.forward + lm_head (:488-619, :720) on (spliced) embeddings; returns (logits, cache, hidden).
        With ``past_key_values`` this is the decode step: ``inputs_embeds`` holds the new token(s) only, every row takes
        the default adapter (``modal_id`` None — the reference drops the modality masks when a cache is present, :436-438)."""
        B, S, H = inputs_embeds.shape
        cfg = self.config
        nH = cfg.num_attention_heads
        cache, past = past_key_values, 0
        if cache is not None:
            past = cache.length
            if past + S > cache.capacity:
                cache.grow(max(past + S, cache.capacity * 2))
        elif use_cache:
            cache = KVCache(len(self.layers), B, S + max(1, int(cache_extra)), nH, H // nH, self.dtype, self.device)
        self._rope_tables(past + S)  # before the workspace: its plans point at the tables
        key = (B, S)
        if key not in self._ws:
            if S > 1:
                for k_ in [k_ for k_ in self._ws if k_[1] > 1]:  # one resident prefill shape: its buffers are GBs at 7B width
                    del self._ws[k_]
            self._ws[key] = _Workspace(self, B, S)
        ws = self._ws[key]
        if ws.pos_value != past:
            ws.pos.fill_(past)
            ws.pos_value = past
        self._last_group_seq = ws.set_routing(modal_id, modal_lut)
        ws.load_rows(inputs_embeds.reshape(B * S, H), ws.x)
        full = attention_mask is None or bool(attention_mask.all())  # one host sync per call, not per layer
        hidden = []
        for li, (layer, plans) in enumerate(zip(self.layers, ws.plans)):
            if output_hidden_states:
                hidden.append(ws.sequence_order(ws.x))
            dense = ws.dense  # materialised form: no rank-space (LoRA-down) launches
            self._rmsnorm(ws.x, layer.ln1, ws.xn)
            if not dense:
                plans["down_qkv"].run()
            plans["up_qkv"].run()
            self._attention(ws, attention_mask, full, cache, li, past)
            if not dense:
                plans["down_o"].run()
            plans["up_o"].run()
            self._rmsnorm(ws.x, layer.ln2, ws.xn)
            if not dense:
                plans["down_gu"].run()
            plans["up_g"].run()
            plans["up_u"].run()
            if not dense:
                plans["down_d"].run()
            plans["up_d"].run()
        if cache is not None:
            cache.length = past + S
        self._rmsnorm(ws.x, self.norm, ws.xn)
        if output_hidden_states:
            hidden.append(ws.sequence_order(ws.xn))
        if last_logits_only and S > 1:  # generate(): only the last position feeds the sampler
            LN.gather_rows(ws.xn, ws.last_rows(), ws.xn_last)
            ws.lm_head_last.run()
            return ws.logits_last.view(B, 1, -1), cache, (tuple(hidden) if output_hidden_states else None)
        ws.lm_head.run()
        return ws.logits.view(B, S, -1), cache, (tuple(hidden) if output_hidden_states else None)

    def _decode_native(self, ids, cache, output_hidden_states) -> bool:
        D = self.config.hidden_size // self.config.num_attention_heads
        return (DECODE_NATIVE and ids.dim() == 2 and ids.shape[1] == 1 and ids.shape[0] <= DC.MAX_M and D == 128
                and not output_hidden_states and cache.length > 0)

    def _decode_workspace(self, cache: "KVCache", attention_mask=None) -> "DC.DecodeWorkspace":
        """The decode step's buffers / launches / graph for this cache (rebuilt when the cache was reallocated).  Padded
        positions of the prompt are taken from ``attention_mask`` (or the mask the prefill stored) when the workspace is
        built; every later position is attended to, as HF generate's appended ones columns have it."""
        dws = self._dws
        if dws is None or not dws.matches(cache) or dws.rope_ptr != self._rope[0].data_ptr():
            mask = attention_mask if attention_mask is not None else getattr(cache, "prefill_mask", None)
            key_mask = None
            if mask is not None and not bool(mask.all()):
                key_mask = torch.ones((mask.shape[0], cache.capacity), dtype=torch.uint8, device=self.device)
                n = min(mask.shape[1], cache.capacity)
                key_mask[:, :n] = (mask[:, :n] != 0).to(torch.uint8)
            self._dws = None  # release the old graph and buffers first
            dws = DC.DecodeWorkspace(self, cache, key_mask, use_graph=DECODE_GRAPH)
            dws.rope_ptr = self._rope[0].data_ptr()
            self._dws = dws
        return dws

    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, inputs_embeds=None, labels=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, modal_inputs=None, return_dict=None,
                last_logits_only: bool = False, cache_extra: int = 128):
        """Reference signature (multimodal_llama.py:676-688).  ``past_key_values`` (a ``KVCache`` returned by an earlier
        call with ``use_cache=True``) selects the decode step: new tokens only, default adapter, no splice (:290-293)."""
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised on this path")
        if past_key_values is not None:
            if not isinstance(past_key_values, KVCache):
                raise TypeError("past_key_values must be the KVCache a previous forward(use_cache=True) returned")
            if input_ids is None:
                raise ValueError("the decode step takes input_ids")
            ids = input_ids.to(self.device).contiguous()
            if attention_mask is not None and modal_inputs is not None and ids.shape[1] == 1:
                # multimodal_arch.py:291-292: the mask is rebuilt as all ones over past + 1
                attention_mask = torch.ones((ids.shape[0], past_key_values.length + 1), dtype=attention_mask.dtype, device=self.device)
            if self._decode_native(ids, past_key_values, output_hidden_states):
                # one token per sequence, batch <= 64: the graph-captured HBM-bound step of modelcompose_b200/decode.py
                cache = past_key_values
                if cache.length + 1 > cache.capacity:
                    cache.grow(max(cache.length + 1, cache.capacity * 2))
                self._rope_tables(cache.length + 1)
                dws = self._decode_workspace(cache, attention_mask)
                logits = dws.step(ids, cache.length).view(ids.shape[0], 1, -1)
                cache.length += 1
                out = CausalLMOutputWithPast(logits=logits, past_key_values=cache)
                return (logits, cache) if return_dict is False else out
            r = SP.splice(ids, None, None, self.embed_tokens, {})  # embedding lookup of the new tokens
            logits, kv, hidden = self.prefill(r.inputs_embeds, None, attention_mask, True, bool(output_hidden_states), past_key_values)
            out = CausalLMOutputWithPast(logits=logits, past_key_values=kv, hidden_states=hidden)
            return (logits, kv) if return_dict is False else out
        modal_id = None
        if input_ids is not None:
            feats = self.project_modal_features(modal_inputs) if modal_inputs else {}
            pre = {m: self.prefix_tokens[m] for m in feats} if self.prefix_tokens is not None else None
            suf = {m: self.suffix_tokens[m] for m in feats} if self.suffix_tokens is not None else None
            r = SP.splice(input_ids.to(self.device).contiguous(),
                          None if attention_mask is None else attention_mask.to(self.device).contiguous(),
                          None if labels is None else labels.to(self.device).contiguous(), self.embed_tokens, feats, pre, suf,
                          list(modal_inputs.keys()) if modal_inputs else None)
            inputs_embeds, attention_mask, labels = r.inputs_embeds, r.attention_mask, r.labels
            modal_lut = None
            if r.modal_names:
                # splice ids follow the order of `feats`; routing ids follow self.modal_names (0 = default): the mapping rides into
                # the permutation kernel as a 16-entry table
                modal_lut = [0] * (SP.MAX_MODAL + 1)
                for i, m in enumerate(r.modal_names):
                    modal_lut[1 + i] = self.modal_names.index(m)
                modal_id = r.modal_id
            if self.config.lora_strategy not in ("modal", "modal+language"):  # :703-704
                modal_id, modal_lut = None, None
        logits, kv, hidden = self.prefill(inputs_embeds, modal_id, attention_mask, bool(use_cache), bool(output_hidden_states),
                                          last_logits_only=last_logits_only and labels is None, cache_extra=cache_extra,
                                          modal_lut=modal_lut if input_ids is not None else None)
        modal_id = self._last_group_seq  # routing groups in sequence order (the splice's ids mapped through the table)
        if kv is not None:
            kv.prefill_mask = attention_mask  # spliced mask of the prompt (padded positions stay masked in the decode steps)
        loss = None
        if labels is not None:  # :723-733
            shift_logits = logits[..., :-1, :].contiguous().view(-1, self.config.vocab_size)
            shift_labels = labels[..., 1:].contiguous().view(-1)
            loss = torch.nn.functional.cross_entropy(shift_logits.float(), shift_labels)
        out = CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=kv, hidden_states=hidden, modal_id=modal_id)
        if return_dict is False:
            t = (logits, kv) if kv is not None else (logits,)
            return (loss,) + t if loss is not None else t
        return out

    __call__ = forward

    @torch.no_grad()
    def generate(self, input_ids, modal_inputs=None, attention_mask=None, max_new_tokens: int = 128, do_sample: bool = False,
                 temperature: float = 1.0, top_p: Optional[float] = None, use_cache: bool = True, eos_token_id: Optional[int] = None,
                 pad_token_id: Optional[int] = None, generator: Optional[torch.Generator] = None, **_):
        """Greedy / nucleus decoding with the call shape of the reference's eval loop
        (``model.generate(input_ids, modal_inputs=..., do_sample=..., temperature=..., top_p=..., max_new_tokens=...,
        use_cache=True)``, modelcompose/eval/model_multimodal_qa_loader.py:93-102).  Returns ``[B, S + new]`` ids: the
        prompt (sentinels included, as HF returns it) followed by the generated tokens."""
        ids = input_ids.to(self.device)
        B = ids.shape[0]
        if attention_mask is None:
            attention_mask = torch.ones_like(ids)
        out = self.forward(ids, attention_mask.to(self.device), modal_inputs=modal_inputs, use_cache=True, last_logits_only=True,
                           cache_extra=max_new_tokens + 1)
        cache = out.past_key_values
        logits = out.logits[:, -1, :]
        text_mask = attention_mask.to(self.device)
        done = torch.zeros(B, dtype=torch.bool, device=self.device)
        pad = pad_token_id if pad_token_id is not None else (eos_token_id if eos_token_id is not None else 0)
        greedy = not (do_sample and temperature > 0)
        D = self.config.hidden_size // self.config.num_attention_heads
        native = DECODE_NATIVE and B <= DC.MAX_M and D == 128
        dws = None
        new_tokens = []
        for step in range(max_new_tokens):
            if not greedy:
                probs = torch.softmax(logits.float() / temperature, dim=-1)
                if top_p is not None and top_p < 1.0:
                    sp, si = probs.sort(dim=-1, descending=True)
                    keep = sp.cumsum(-1) - sp < top_p
                    sp = sp * keep
                    probs = torch.zeros_like(probs).scatter_(1, si, sp)
                    probs = probs / probs.sum(-1, keepdim=True)
                nxt = torch.multinomial(probs, 1, generator=generator).squeeze(1)
            elif dws is not None:
                nxt = dws.next64.clone()  # the step's own argmax (mc_argmax_rows inside the graph)
            else:
                nxt = torch.empty(B, dtype=torch.int64, device=self.device)
                DC.argmax_rows(logits.contiguous(), None, nxt)
            if eos_token_id is not None:
                nxt = torch.where(done, torch.full_like(nxt, pad), nxt)
            new_tokens.append(nxt)
            if eos_token_id is not None:
                done = done | (nxt == eos_token_id)
                if bool(done.all()):
                    break
            if step + 1 == max_new_tokens:
                break
            if native:
                # graph-captured decode step (modelcompose_b200/decode.py): with modal inputs every cached position is attended
                # to (multimodal_arch.py:291-292); text-only batches keep the prompt's padded positions masked (HF generate
                # appends ones to the caller's mask).  The graph's own argmax already wrote the next ids and advanced the
                # position; only sampled / eos-padded tokens have to be written over them.
                if dws is None:
                    self._rope_tables(cache.capacity)
                    if modal_inputs:
                        cache.prefill_mask = None
                    dws = self._decode_workspace(cache)
                    dws.ids.copy_(nxt)
                    dws.pos.fill_(cache.length)
                elif not greedy or eos_token_id is not None:
                    dws.ids.copy_(nxt)
                dws.run()
                cache.length += 1
                logits = dws.logits
                continue
            if modal_inputs:
                # multimodal_arch.py:291-292: with modal inputs the mask is rebuilt as all ones over past + 1
                step_mask = torch.ones((B, cache.length + 1), dtype=attention_mask.dtype, device=self.device)
            else:
                # text-only: HF generate appends a ones column to the caller's mask, so padded positions stay masked
                text_mask = torch.cat([text_mask, torch.ones((B, 1), dtype=text_mask.dtype, device=self.device)], dim=1)
                step_mask = text_mask
            o = self.forward(nxt[:, None], step_mask, past_key_values=cache, use_cache=True)
            logits = o.logits[:, -1, :]
        return torch.cat([ids, torch.stack(new_tokens, dim=1)], dim=1)
