"""Modality-token splice: host mirror of the reference ``prepare_inputs_labels_for_multimodal``.

Mirrors ``modelcompose/model/multimodal_arch.py:287-459`` of the reference (same return tuple, mask dtypes,
ragged-batch behaviour and failure modes).  All data movement is done by
``mc_splice_*`` in the CUDA library (include/modelcompose_b200.h) — one scan, one descriptor and one gather
launch per batch instead of the reference's per-sample Python loop; there is no torch/CPU path here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi

IGNORE_INDEX = -100  # modelcompose/constants.py:7
# modelcompose/constants.py:23-30
MODAL_TOKEN_INDEXES = {"vision": -200, "relrep": -201, "text": -202, "audio": -203, "video": -204, "point": -205}
MAX_MODAL = 6


class SpliceModal(C.Structure):
    _fields_ = [("sentinel", C.c_int64), ("n_blocks", C.c_int32), ("n_rows", C.c_int32), ("n_prefix", C.c_int32),
                ("n_suffix", C.c_int32), ("features", C.c_void_p), ("prefix", C.c_void_p), ("suffix", C.c_void_p),
                ("mask_out", C.c_void_p)]


class SpliceIO(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("hidden", C.c_int32), ("embed_table", C.c_void_p),
                ("attention_mask_in", C.c_void_p), ("mask_elem_size", C.c_int32), ("labels_in", C.c_void_p),
                ("out_embeds", C.c_void_p), ("out_modal_id", C.c_void_p), ("out_attention_mask", C.c_void_p),
                ("out_labels", C.c_void_p), ("out_default_mask", C.c_void_p)]


@dataclass
class SpliceResult:
    inputs_embeds: torch.Tensor                 # [B, S', H]
    modal_id: torch.Tensor                      # uint8 [B, S']: 0 = default, 1 + i = modal_names[i]
    attention_mask: Optional[torch.Tensor]      # [B, S'] in the input mask dtype
    labels: Optional[torch.Tensor]              # int64 [B, S']
    modal_attention_mask: Optional[Dict[str, torch.Tensor]]
    out_len: List[int]                          # per-sample length before padding
    modal_names: List[str]
    algorithmic_bytes: int


# Reusable plans, keyed by device and batch geometry (a plan owns a few KB..MB of device tables and allocates nothing
# in steady state).  A handful of shapes are live at a time (one prefill shape, one decode shape).
_PLANS: "OrderedDict[tuple, C.c_void_p]" = None
_MAX_PLANS = 8


def _plan_for(device, B, S, V, modals, n_modal) -> C.c_void_p:
    global _PLANS
    from collections import OrderedDict
    if _PLANS is None:
        _PLANS = OrderedDict()
    key = (device.index, B, S, V, tuple((int(modals[i].sentinel), modals[i].n_blocks, modals[i].n_rows, modals[i].n_prefix,
                                          modals[i].n_suffix) for i in range(n_modal)))
    plan = _PLANS.get(key)
    if plan is None:
        plan = C.c_void_p()
        _cabi.check(_cabi.lib().mc_splice_plan_create(C.byref(plan), B, S, V, modals, n_modal), "mc_splice_plan_create")
        _PLANS[key] = plan
        while len(_PLANS) > _MAX_PLANS:
            _, old = _PLANS.popitem(last=False)
            _cabi.lib().mc_splice_plan_destroy(old)
    else:
        _PLANS.move_to_end(key)
    return plan


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check(t: torch.Tensor, what: str, dtype=None):
    if not t.is_cuda:
        raise ValueError(f"{what} must be a CUDA tensor (modelcompose_b200 has no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{what}: expected {dtype}, got {t.dtype}")


def splice(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], labels: Optional[torch.Tensor],
           embed_table: torch.Tensor, modal_features: Dict[str, torch.Tensor],
           prefix_tokens: Optional[Dict[str, torch.Tensor]] = None,
           suffix_tokens: Optional[Dict[str, torch.Tensor]] = None,
           modal_input_keys: Optional[Sequence[str]] = None) -> SpliceResult:
    """``modal_features``: modality → [n_blocks, n_m, H] projected features in ``infer_modals`` order, WITHOUT
    prefix/suffix (the kernel splices ``prefix_tokens[m]`` / ``suffix_tokens[m]`` ([1, p, H]) around each block,
    which is what ``encode_modal_inputs`` :244-253 concatenates)."""
    lib = _cabi.lib()
    names = list(modal_features.keys())
    if len(names) > MAX_MODAL:
        raise ValueError(f"at most {MAX_MODAL} modalities")
    _check(input_ids, "input_ids", torch.int64)
    _check(embed_table, "embed_table")
    B, S = input_ids.shape
    V, H = embed_table.shape
    dtype = embed_table.dtype
    mask_dtype = attention_mask.dtype if attention_mask is not None else torch.bool
    if mask_dtype not in (torch.bool, torch.int64):
        raise ValueError(f"attention_mask dtype {mask_dtype} unsupported (bool or int64, as the reference passes)")
    elem = 1 if mask_dtype == torch.bool else 8
    if attention_mask is not None:
        _check(attention_mask, "attention_mask")
    if labels is not None:
        _check(labels, "labels", torch.int64)

    modals = (SpliceModal * max(len(names), 1))()
    keep = []
    for i, m in enumerate(names):
        f = modal_features[m]
        _check(f, f"modal_features[{m}]", dtype)
        if f.dim() != 3 or f.shape[2] != H:
            raise ValueError(f"modal_features[{m}] must be [n_blocks, n_rows, {H}]")
        pre = prefix_tokens[m] if prefix_tokens is not None and m in prefix_tokens else None
        suf = suffix_tokens[m] if suffix_tokens is not None and m in suffix_tokens else None
        for t, w in ((pre, "prefix"), (suf, "suffix")):
            if t is not None:
                _check(t, f"{w}_tokens[{m}]", dtype)
        keep += [f, pre, suf]
        modals[i].sentinel = MODAL_TOKEN_INDEXES[m]
        modals[i].n_blocks, modals[i].n_rows = f.shape[0], f.shape[1]
        modals[i].n_prefix = 0 if pre is None else pre.shape[-2]
        modals[i].n_suffix = 0 if suf is None else suf.shape[-2]
        modals[i].features, modals[i].prefix, modals[i].suffix = _ptr(f), _ptr(pre), _ptr(suf)

    stream = _cabi.current_stream_ptr()
    plan = _plan_for(input_ids.device, B, S, V, modals, len(names))
    _cabi.check(lib.mc_splice_plan_scan(plan, input_ids.data_ptr(), stream), "mc_splice_plan_scan")
    max_len, min_len = C.c_int(), C.c_int()
    out_len = (C.c_int32 * B)()
    used = (C.c_int32 * MAX_MODAL)()
    _cabi.check(lib.mc_splice_plan_info(plan, C.byref(max_len), C.byref(min_len), out_len, used), "mc_splice_plan_info")
    Sp = max_len.value
    if min_len.value != Sp and labels is None:
        # the reference binds `_new_labels` only under `if labels is not None` (multimodal_arch.py:414-429)
        raise UnboundLocalError("cannot access local variable '_new_labels' where it is not associated with a value "
                                "(ragged batch with labels=None, reference multimodal_arch.py:414-429)")
    any_sentinel = any(used[i] > 0 for i in range(len(names)))
    mask_names = list(names)
    if modal_input_keys is not None and any_sentinel:
        mask_names = [m for m in names if m in set(modal_input_keys)]
        if len(mask_names) != len(names) and any(n == S for n in out_len):
            raise ValueError("a batch mixing sentinel-free samples with a partial modal_inputs dict has no "
                             "consistent mask shape in the reference (multimodal_arch.py:323-342,452)")
    dev = input_ids.device
    embeds = torch.empty((B, Sp, H), dtype=dtype, device=dev)
    modal_id = torch.empty((B, Sp), dtype=torch.uint8, device=dev)
    attn_out = torch.empty((B, Sp), dtype=mask_dtype, device=dev) if attention_mask is not None else None
    labels_out = torch.empty((B, Sp), dtype=torch.int64, device=dev) if labels is not None else None
    masks = {m: torch.empty((B, Sp), dtype=mask_dtype, device=dev) for m in mask_names}
    default_mask = torch.empty((B, Sp), dtype=torch.bool, device=dev) if masks else None
    for i, m in enumerate(names):
        modals[i].mask_out = _ptr(masks.get(m))
    io = SpliceIO(_cabi.dtype_code(dtype), H, embed_table.data_ptr(), _ptr(attention_mask), elem, _ptr(labels),
                  embeds.data_ptr(), modal_id.data_ptr(), _ptr(attn_out), _ptr(labels_out), _ptr(default_mask))
    _cabi.check(lib.mc_splice_run(plan, C.byref(io), modals, stream), "mc_splice_run")
    _cabi.count_launch(3)  # scan + expand (plan) + gather
    nbytes = int(lib.mc_splice_plan_bytes(plan, H * embed_table.element_size()))
    del keep
    out_masks = None
    if masks:
        out_masks = dict(masks)
        out_masks["default"] = default_mask  # (Σ modal masks == 0), multimodal_arch.py:452-453
    return SpliceResult(embeds, modal_id, attn_out, labels_out, out_masks, list(out_len), names, nbytes)


def prepare_inputs_labels_for_multimodal(input_ids, attention_mask, past_key_values, labels, modal_features,
                                         prefix_tokens, suffix_tokens, embed_table, modal_input_keys=None):
    """Reference signature (:287-289) with the encoder/projector outputs passed in place of raw ``modal_inputs``.
    Returns the reference 6-tuple ``(None, attention_mask, past_key_values, inputs_embeds, labels, masks)``."""
    if modal_features is None or input_ids.shape[1] == 1:  # decode step early-return (:290-293)
        if past_key_values is not None and modal_features is not None and input_ids.shape[1] == 1:
            attention_mask = torch.ones((attention_mask.shape[0], past_key_values[-1][-1].shape[-2] + 1),
                                        dtype=attention_mask.dtype, device=attention_mask.device)
        return input_ids, attention_mask, past_key_values, None, labels, None
    r = splice(input_ids, attention_mask, labels, embed_table, modal_features, prefix_tokens, suffix_tokens,
               modal_input_keys)
    return None, r.attention_mask, past_key_values, r.inputs_embeds, r.labels, r.modal_attention_mask
