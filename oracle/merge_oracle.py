"""Oracle for the merge path.  TEST INFRASTRUCTURE.

Follows ``scripts/model_composition/merge_unimodal_modelcompose.py`` of the reference;
line numbers below are into that file unless another file is named.
"""
from __future__ import annotations

import json
import os
from collections import defaultdict
from typing import Dict, List, Sequence

import torch

# merge_unimodal_modelcompose.py:15-21 (dict order matters: first matching key wins)
MODAL_DICT = {
    "mm_vision_encoder": "vision",
    "mm_vision_tower": "vision",
    "mm_vision2_encoder": "vision2",
    "mm_vision2_tower": "vision2",
    "mm_video_encoder": "video",
    "mm_audio_encoder": "audio",
    "mm_point_encoder": "point",
}


def get_modal_from_config(config: dict) -> str:
    """:22-26 — modality of a checkpoint = first MODAL_DICT key present as non-empty str."""
    for key in MODAL_DICT:
        if key in config.keys() and isinstance(config[key], str) and len(config[key]) > 0:
            return MODAL_DICT[key]
    assert False, "No modality is recognized, please check the config."


def group_weights(state_dicts: Sequence[Dict[str, torch.Tensor]]) -> Dict[str, List[torch.Tensor]]:
    """:30-40 — group tensors by key, first-seen key order across inputs."""
    weights_to_merge = defaultdict(list)
    for sd in state_dicts:
        for key in sd:
            weights_to_merge[key].append(sd[key])
    return weights_to_merge


def merge_weights(state_dicts, configs, strategy: str) -> Dict[str, torch.Tensor]:
    """:94-115 — the ``online-merge-*`` rename branch and the literal ``sum`` / ``mean`` branch."""
    weights_to_merge = group_weights(state_dicts)
    if strategy.startswith("online-merge-"):
        merged = dict()
        modal_names = [get_modal_from_config(c) for c in configs]
        for key in weights_to_merge:
            if len(weights_to_merge[key]) == 1:
                merged[key] = weights_to_merge[key][0]
            else:
                assert "default" in key
                for modal_name, weight in zip(modal_names, weights_to_merge[key]):
                    merged[key.replace("default", f"default-{modal_name}")] = weight
        return merged
    if strategy == "sum":
        return {k: ref_sum(v) for k, v in weights_to_merge.items()}
    if strategy == "mean":
        return {k: ref_sum(v) / len(v) for k, v in weights_to_merge.items()}
    raise NotImplementedError(strategy)


def ref_sum(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """:108,112 — Python ``sum(list)``: ((0 + t0) + t1) + ... each add rounded in the storage dtype."""
    acc = 0
    for t in tensors:
        acc = acc + t
    return acc


def merge_configs(configs: Sequence[dict], strategy: str):
    """:117-136 — first-truthy union; the strategy string is consumed on the FIRST config only.

    Returns (merged_configs, strategy_after) — ``strategy_after`` is what lands in merge_info.txt (:144)."""
    merged = {}
    for config in configs:
        for key in config:
            if key in merged:
                merged[key] = merged[key] or config[key]
            else:
                merged[key] = config[key]
        if strategy.startswith("online-merge-"):
            strategy = strategy.replace("online-merge-", "")
            if strategy.startswith("reset-"):
                merged["reset_scaling_weights"] = strategy.replace("reset-", "")
            else:
                merged["merge_default_weights"] = strategy
    for config in configs:
        modal_name = get_modal_from_config(config)
        merged[f"{modal_name}_lora_alpha"] = config["lora_alpha"]
        merged[f"{modal_name}_lora_r"] = config["lora_r"]
    return merged, strategy


def merge_info_text(filepaths: Sequence[str], strategy_after: str, output_path: str) -> str:
    """:142-144."""
    inputs = "\n".join(filepaths)
    return f"Inputs:\n{inputs}\n\nOutput({strategy_after}):{output_path}"


def merge_checkpoint_dirs(filepaths: Sequence[str], output_path: str, strategy: str):
    """:28-145 end to end on directories (load → merge → save); returns (weights, config)."""
    sds, cfgs = [], []
    for fp in filepaths:
        ap = os.path.join(fp, "adapter_model.bin")
        if not os.path.exists(ap):
            ap = os.path.join(fp, "mm_projector.bin")
        sds.append(torch.load(ap, map_location="cpu"))
        cfgs.append(json.load(open(os.path.join(fp, "config.json"))))
    merged = merge_weights(sds, cfgs, strategy)
    mcfg, strategy_after = merge_configs(cfgs, strategy)
    os.makedirs(output_path, exist_ok=True)
    torch.save(merged, os.path.join(output_path, "adapter_model.bin"))
    json.dump(mcfg, open(os.path.join(output_path, "config.json"), "w"), indent=4)
    with open(os.path.join(output_path, "merge_info.txt"), "w") as f:
        f.write(merge_info_text(filepaths, strategy_after, output_path))
    return merged, mcfg


# ----------------------------------------------------------------------------- A8: coefficients
def extract_params(input_string: str) -> Dict[str, float]:
    """multimodal_llama.py:109-118 — "k=v,k=v" → {k: float(v)}."""
    params = {}
    for pair in input_string.split(","):
        key, value = pair.split("=")
        params[key.strip()] = float(value)
    return params


def effective_scaling(modal_names: Sequence[str], r: int, lora_alpha: float, reset_scaling_weights):
    """multimodal_llama.py:84-106 — adapter list + scaling dict after the reset coefficients.

    Returns (adapter_names, scaling, default_adapter_names or None).  Python float64 arithmetic."""
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


# ----------------------------------------------------------------------------- the N-source merge
def weighted_merge(tensors: Sequence[torch.Tensor], weights: Sequence[float], out_dtype=None) -> torch.Tensor:
    """The 7B-shaped N-source elementwise merge:
    ``out = ((w0*t0 + w1*t1) + w2*t2) ...`` with fp32 products and fp32 left-to-right adds
    (separate mul and add — no FMA), one round-to-nearest-even to ``out_dtype`` at the end.
    Weights are rounded to fp32 first, as the kernel receives them."""
    out_dtype = out_dtype or tensors[0].dtype
    acc = None
    for t, w in zip(tensors, weights):
        term = t.to(torch.float32) * torch.tensor(float(w), dtype=torch.float32)
        acc = term if acc is None else acc + term
    return acc.to(out_dtype)


def delta_weight(A: torch.Tensor, B: torch.Tensor, scale: float) -> torch.Tensor:
    """scripts/model_composition/delta_weights_compare.py:24-31 — ``(B @ A) * scale``."""
    return (B.to(torch.float32) @ A.to(torch.float32)) * scale


def materialise_effective_weight(W, As, Bs, scales, out_dtype=None) -> torch.Tensor:
    """``W_eff = W + Σ_m s_m · B_m @ A_m`` (delta_weights_compare.py:61; convert_to_multimodal.py:111-113),
    fp32 accumulation, one rounding."""
    out_dtype = out_dtype or W.dtype
    acc = W.to(torch.float32)
    for A, B, s in zip(As, Bs, scales):
        acc = acc + delta_weight(A, B, s)
    return acc.to(out_dtype)
