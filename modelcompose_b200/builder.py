"""``load_pretrained_model`` — drop-in for the multimodal branch of the reference loader.

Mirrors ``modelcompose/model/builder.py:27,138-185,223-231``: the merged ``config.json`` (with
``reset_scaling_weights``) configures the model, base LLM weights come from ``model_base``, every adapter starts
from the reset init (B = 0) and ``adapter_model.bin`` (fallback ``mm_projector.bin``) plus the optional
``non_lora_trainables.bin`` are loaded non-strictly and cast to fp16; the model lands on ``device`` in fp16.
Weights go straight into the packed layouts of the routed kernels (``linear.pack_adapters``).  Only the
``'multimodal' in model_name`` branch exists here — the LLaVA / MPT / PEFT branches are outside the hot path
(SURVEY.md §2 rows 9-11).  The frozen modality encoders are not loaded (out of scope): ``modal_processors`` is
``None`` and ``modal_inputs`` must carry encoder features.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict

import torch

from .model import MultimodalConfig, MultimodalLlamaForCausalLM


def _load_base_state_dict(model_base: str) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    st = sorted(glob.glob(os.path.join(model_base, "*.safetensors")))
    if st:
        from safetensors.torch import load_file
        for f in st:
            sd.update(load_file(f, device="cpu"))
        return sd
    bins = sorted(glob.glob(os.path.join(model_base, "pytorch_model*.bin")))
    if not bins:
        raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {model_base}")
    for f in bins:
        sd.update(torch.load(f, map_location="cpu"))
    return sd


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda", torch_dtype=torch.float16, materialize=None, decode_dense=None):
    """Returns ``(tokenizer, model, modal_processors, context_len)`` like the reference (builder.py:231).
    ``materialize`` (extra, default: the MC_MATERIALIZE environment switch): build one dense effective weight per routing
    group at load (``modelcompose_b200.materialize``) and run every linear as a grouped GEMM instead of base + LoRA branches.
    ``decode_dense`` (extra, default: MC_DECODE_DENSE): keep the branch form for the prefill and build the dense weight of the text
    group only, for the decode steps (model.py: DECODE_DENSE)."""
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes quantised loading is outside the B200 hot path (SURVEY.md §2.2)")
    if "multimodal" not in model_name.lower():
        raise NotImplementedError("only the 'multimodal' branch of the reference loader is implemented "
                                  "(the checkpoint directory's basename must contain 'multimodal', README.md:96)")
    with open(os.path.join(model_path, "config.json")) as f:
        cfg = MultimodalConfig.from_dict(json.load(f))
    tokenizer = None
    tok_dir = model_base if model_base is not None else model_path
    if os.path.exists(os.path.join(tok_dir, "tokenizer.model")) or os.path.exists(os.path.join(tok_dir, "tokenizer.json")):
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(tok_dir, use_fast=False)
    base = _load_base_state_dict(model_base if model_base is not None else model_path)
    adapters: Dict[str, torch.Tensor] = {}
    if model_base is not None:
        adapter_path = os.path.join(model_path, "adapter_model.bin")
        if not os.path.exists(adapter_path):
            adapter_path = os.path.join(model_path, "mm_projector.bin")
        adapters = {k: v.to(torch.float16) for k, v in torch.load(adapter_path, map_location="cpu").items()}
        extra = os.path.join(model_path, "non_lora_trainables.bin")
        if os.path.exists(extra):
            adapters.update({k: v.to(torch.float16) for k, v in torch.load(extra, map_location="cpu").items()})
    else:  # merged weights and adapters saved together (builder.py:169-180)
        adapters = base
    model = MultimodalLlamaForCausalLM(cfg, base, adapters, device=device, dtype=torch_dtype, materialize=materialize,
                                       decode_dense=decode_dense)
    context_len = getattr(cfg, "max_sequence_length", 2048)
    return tokenizer, model, None, context_len
