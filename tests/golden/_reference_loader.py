"""Import the UNMODIFIED reference (/root/reference) in the authoring container.

Test infrastructure only.  Used by ``make_golden.py`` (fixture generation) and by
the ``-m "not gpu"`` tests that cross-check ``oracle/`` against the live
reference *when /root/reference exists* (it does not exist on the GPU box; those
tests skip there and the committed fixtures take over).

The reference pins ``peft==0.4.0`` and ``transformers==4.31.0`` (pyproject.toml:17-18);
this container has neither peft nor that transformers.  The shims below restate the
few third-party symbols the hot path touches:

* ``peft.tuners.lora.{LoraLayer,Linear}`` / ``peft.utils.transpose`` (peft 0.4.0),
* ``LlamaRotaryEmbedding`` / ``apply_rotary_pos_emb`` / ``rotate_half`` with the
  transformers-4.31 signatures (5.5 changed them),
* the names 4.31's ``from modeling_llama import *`` exported.

Everything else (merge CLI, LocalLoraLinear/Attention/MLP/DecoderLayer,
prepare_inputs_labels_for_multimodal, build_vision_projector, constants) is
the reference's own code, executed from where it lies.
"""
from __future__ import annotations

import builtins
import importlib.util
import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MODELCOMPOSE_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scripts", "model_composition",
                                       "merge_unimodal_modelcompose.py"))


# --------------------------------------------------------------------------- peft 0.4.0 shim
def _transpose(weight, fan_in_fan_out):
    return weight.T if fan_in_fan_out else weight


class _LoraLayer:
    """peft==0.4.0 ``peft.tuners.lora.LoraLayer`` (semantics the reference uses)."""

    def __init__(self, in_features: int, out_features: int, **kwargs):
        self.r = {}
        self.lora_alpha = {}
        self.scaling = {}
        self.lora_dropout = nn.ModuleDict({})
        self.lora_A = nn.ModuleDict({})
        self.lora_B = nn.ModuleDict({})
        self.lora_embedding_A = nn.ParameterDict({})
        self.lora_embedding_B = nn.ParameterDict({})
        self.merged = False
        self.disable_adapters = False
        self.in_features = in_features
        self.out_features = out_features
        self.kwargs = kwargs

    def update_layer(self, adapter_name, r, lora_alpha, lora_dropout, init_lora_weights):
        self.r[adapter_name] = r
        self.lora_alpha[adapter_name] = lora_alpha
        drop = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else nn.Identity()
        self.lora_dropout.update(nn.ModuleDict({adapter_name: drop}))
        if r > 0:
            self.lora_A.update(nn.ModuleDict({adapter_name: nn.Linear(self.in_features, r, bias=False)}))
            self.lora_B.update(nn.ModuleDict({adapter_name: nn.Linear(r, self.out_features, bias=False)}))
            self.scaling[adapter_name] = lora_alpha / r
        if init_lora_weights:
            self.reset_lora_parameters(adapter_name)
        self.to(self.weight.device)

    def reset_lora_parameters(self, adapter_name):
        if adapter_name in self.lora_A.keys():
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B[adapter_name].weight)


class _LoraLinear(nn.Linear, _LoraLayer):
    """peft==0.4.0 ``peft.tuners.lora.Linear.__init__`` (forward is overridden by the reference)."""

    def __init__(self, adapter_name, in_features, out_features, r=0, lora_alpha=1, lora_dropout=0.0,
                 fan_in_fan_out=False, is_target_conv_1d_layer=False, **kwargs):
        init_lora_weights = kwargs.pop("init_lora_weights", True)
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        _LoraLayer.__init__(self, in_features=in_features, out_features=out_features)
        self.weight.requires_grad = False
        self.fan_in_fan_out = fan_in_fan_out
        if fan_in_fan_out:
            self.weight.data = self.weight.data.T
        nn.Linear.reset_parameters(self)
        self.update_layer(adapter_name, r, lora_alpha, lora_dropout, init_lora_weights)
        self.active_adapter = adapter_name
        self.is_target_conv_1d_layer = is_target_conv_1d_layer


# --------------------------------------------------------------------------- transformers 4.31 shims
class _LlamaRotaryEmbedding(nn.Module):
    """transformers==4.31.0 ``LlamaRotaryEmbedding`` (cached cos/sin, forward(x, seq_len))."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, device=None):
        super().__init__()
        self.dim = dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2).float().to(device) / self.dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._set_cos_sin_cache(max_position_embeddings, self.inv_freq.device, torch.get_default_dtype())

    def _set_cos_sin_cache(self, seq_len, device, dtype):
        self.max_seq_len_cached = seq_len
        t = torch.arange(self.max_seq_len_cached, device=device, dtype=self.inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        self.register_buffer("cos_cached", emb.cos()[None, None, :, :].to(dtype), persistent=False)
        self.register_buffer("sin_cached", emb.sin()[None, None, :, :].to(dtype), persistent=False)

    def forward(self, x, seq_len=None):
        if seq_len > self.max_seq_len_cached:
            self._set_cos_sin_cache(seq_len, x.device, x.dtype)
        return (self.cos_cached[:, :, :seq_len, ...].to(dtype=x.dtype),
                self.sin_cached[:, :, :seq_len, ...].to(dtype=x.dtype))


def _rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def _apply_rotary_pos_emb(q, k, cos, sin, position_ids):
    cos = cos.squeeze(1).squeeze(0)
    sin = sin.squeeze(1).squeeze(0)
    cos = cos[position_ids].unsqueeze(1)
    sin = sin[position_ids].unsqueeze(1)
    return (q * cos) + (_rotate_half(q) * sin), (k * cos) + (_rotate_half(k) * sin)


def infer_modals_restated(model_args):
    """Restated ``infer_modals`` (multimodal_encoder/builder.py:119-130) for the package shell
    (the real builder imports every vendored encoder and cannot be imported here)."""
    modals = ["default"]
    if getattr(model_args, "mm_audio_encoder", None) is not None:
        modals.append("audio")
    if getattr(model_args, "mm_vision_encoder", None) is not None or getattr(model_args, "mm_vision_tower", None) is not None:
        modals.append("vision")
    if getattr(model_args, "mm_video_encoder", None):
        modals.append("video")
    if getattr(model_args, "mm_point_encoder", None):
        modals.append("point")
    return modals


_LOADED: dict = {}


def _shell(name: str, path: str | None = None, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    if path is not None:
        mod.__path__ = [path]
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _install_shells():
    if "shells" in _LOADED:
        return
    ref = REFERENCE_ROOT
    # matplotlib is imported (unused) by ties_merging.py:14
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            _shell("matplotlib", pyplot=None)
            _shell("matplotlib.pyplot")
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # peft 0.4.0
    if "peft" not in sys.modules:
        lora = _shell("peft.tuners.lora", LoraLayer=_LoraLayer, Linear=_LoraLinear)
        tuners = _shell("peft.tuners", path="", lora=lora)
        utils = _shell("peft.utils", transpose=_transpose)
        _shell("peft", path="", tuners=tuners, utils=utils)
    # package shells so modelcompose/model/__init__.py (imports every encoder) never runs
    _shell("modelcompose", path=os.path.join(ref, "modelcompose"))
    _shell("modelcompose.model", path=os.path.join(ref, "modelcompose", "model"))
    _shell("modelcompose.model.language_model", path=os.path.join(ref, "modelcompose", "model", "language_model"))
    enc_b = _shell("modelcompose.model.multimodal_encoder.builder", build_modal_encoders=None,
                   infer_modals=infer_modals_restated)
    _shell("modelcompose.model.multimodal_encoder", path="", builder=enc_b)
    _LOADED["shells"] = True


def _exec(name: str, relpath: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_constants():
    _install_shells()
    if "constants" not in _LOADED:
        _LOADED["constants"] = _exec("modelcompose.constants", "modelcompose/constants.py")
    return _LOADED["constants"]


def load_projector_builder():
    """The real ``modelcompose/model/multimodal_projector/builder.py`` (A10)."""
    _install_shells()
    if "proj" not in _LOADED:
        _shell("modelcompose.model.multimodal_projector", path=os.path.join(
            REFERENCE_ROOT, "modelcompose", "model", "multimodal_projector"))
        # Qformer.py (vendored LAVIS BERT, OUT OF SCOPE: SURVEY §2 row 8) does not import under
        # transformers 5.5; builder.py only needs the two names to exist at import time.
        _shell("modelcompose.model.multimodal_projector.Qformer", BertConfig=None, BertLMHeadModel=None)
        _LOADED["proj"] = _exec("modelcompose.model.multimodal_projector.builder",
                                "modelcompose/model/multimodal_projector/builder.py")
        sys.modules["modelcompose.model.multimodal_projector"].builder = _LOADED["proj"]
    return _LOADED["proj"]


def load_arch():
    """The real ``modelcompose/model/multimodal_arch.py`` (A11-A13)."""
    load_constants()
    load_projector_builder()
    if "arch" not in _LOADED:
        _LOADED["arch"] = _exec("modelcompose.model.multimodal_arch", "modelcompose/model/multimodal_arch.py")
    return _LOADED["arch"]


def load_llama():
    """The real ``modelcompose/model/language_model/multimodal_llama.py`` (A6, A8, A9, A14-A16)."""
    load_arch()
    if "llama" not in _LOADED:
        from transformers.modeling_outputs import BaseModelOutputWithPast
        from transformers.activations import ACT2FN
        from transformers.models.llama.modeling_llama import LlamaRMSNorm, repeat_kv
        import logging
        for k, v in dict(BaseModelOutputWithPast=BaseModelOutputWithPast, ACT2FN=ACT2FN, LlamaRMSNorm=LlamaRMSNorm,
                         repeat_kv=repeat_kv, logger=logging.getLogger("ref"), math=math,
                         LlamaRotaryEmbedding=_LlamaRotaryEmbedding, apply_rotary_pos_emb=_apply_rotary_pos_emb,
                         rotate_half=_rotate_half).items():
            setattr(builtins, k, v)
        _LOADED["llama"] = _exec("modelcompose.model.language_model.multimodal_llama",
                                 "modelcompose/model/language_model/multimodal_llama.py")
    return _LOADED["llama"]


def run_merge_cli(argv: list[str]) -> None:
    """Run the reference merge CLI unmodified (merge_unimodal_modelcompose.py:151-162)."""
    import runpy
    _install_shells()
    script_dir = os.path.join(REFERENCE_ROOT, "scripts", "model_composition")
    old_argv, old_path = sys.argv, list(sys.path)
    sys.path.insert(0, script_dir)
    sys.argv = ["merge_unimodal_modelcompose.py"] + list(argv)
    try:
        runpy.run_path(os.path.join(script_dir, "merge_unimodal_modelcompose.py"), run_name="__main__")
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
        for m in ("ties_merging", "calculate_metrics"):
            sys.modules.pop(m, None)


def make_config(**kw):
    """A reference ``MultimodalConfig`` usable with transformers 5.5 (rope_scaling=None, tp=1)."""
    llama = load_llama()
    cfg = llama.MultimodalConfig(**kw)
    cfg.rope_scaling = None
    cfg.pretraining_tp = 1
    return cfg


class SpliceHost(nn.Module):
    """Minimal host for the reference's real ``MultimodalMetaForCausalLM`` methods (A11-A13):
    fake encoders return pre-generated feature tensors, projectors are real modules."""

    def __init__(self, config, embed_tokens: nn.Embedding, projectors: dict, device="cpu"):
        super().__init__()
        self.config = config
        self.embed_tokens = embed_tokens
        self.projectors = nn.ModuleDict(projectors)
        self._device = torch.device(device)

    @property
    def device(self):
        return self._device

    # the "model" object the mixin asks for
    def get_model(self):
        return self

    def get_modal_encoder(self, modal):
        if modal == "audio":
            return lambda **kw: (kw["audio_inputs"], None)
        return lambda x: x

    def get_modal_projector(self, modal):
        return self.projectors[modal]


def make_splice_host(config, embed_tokens, projectors, device="cpu"):
    arch = load_arch()

    class _Host(SpliceHost, arch.MultimodalMetaForCausalLM):
        pass

    return _Host(config, embed_tokens, projectors, device)
