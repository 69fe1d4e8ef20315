"""Oracle for the modality-token splice.  TEST INFRASTRUCTURE.

Follows ``modelcompose/model/multimodal_arch.py`` of the reference (line numbers into that
file) and ``modelcompose/constants.py:23-30``.  Integer work: results must be bit-exact.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

IGNORE_INDEX = -100  # constants.py:7
# constants.py:23-30 (dict order = tie-break order in modal_token_match; ties cannot occur)
MODAL_TOKEN_INDEXES = {"vision": -200, "relrep": -201, "text": -202, "audio": -203, "video": -204, "point": -205}
SEARCH_LIMIT = 10000  # multimodal_arch.py:278


def modal_token_match(ids: np.ndarray):
    """:270-285 — earliest sentinel position (strictly below 10000), else (None, 10000)."""
    modal, start = None, SEARCH_LIMIT
    for name, tok in MODAL_TOKEN_INDEXES.items():
        hits = np.nonzero(ids == tok)[0]
        if hits.size and hits[0] < start:
            start, modal = int(hits[0]), name
    return modal, start


def add_prefix_suffix(features: Dict[str, torch.Tensor], prefix_tokens=None, suffix_tokens=None):
    """:244-253 — cat([prefix.expand(b), feats, suffix.expand(b)], dim=1) per modality."""
    out = {}
    for modal, f in features.items():
        b = f.shape[0]
        parts = []
        if prefix_tokens is not None and modal in prefix_tokens:
            parts.append(prefix_tokens[modal].expand(b, -1, -1))
        parts.append(f)
        if suffix_tokens is not None and modal in suffix_tokens:
            parts.append(suffix_tokens[modal].expand(b, -1, -1))
        out[modal] = torch.cat(parts, dim=1)
    return out


def splice(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], labels: Optional[torch.Tensor],
           embed_table: torch.Tensor, modal_features: Dict[str, torch.Tensor], modal_input_keys=None):
    """:287-459 — ``prepare_inputs_labels_for_multimodal`` after ``encode_modal_inputs``.

    ``modal_features``: modality → [n_blocks, n_m, H] (already projected, prefix/suffix attached),
    in ``infer_modals`` order.  ``modal_input_keys``: the keys of the caller's ``modal_inputs`` dict
    (:320-321 — only those modalities get a per-sample mask list); defaults to all of ``modal_features``.

    Returns (attention_mask, inputs_embeds, labels, modal_attention_mask) — the non-None members of
    the reference 6-tuple.  Ragged batches with ``labels is None`` raise like the reference (:414-429)."""
    ids_np = input_ids.cpu().numpy()
    B, S = ids_np.shape
    H = embed_table.shape[1]
    mask_dtype = attention_mask.dtype if attention_mask is not None else torch.bool
    modal_input_keys = list(modal_input_keys) if modal_input_keys is not None else list(modal_features.keys())
    cursor = {m: 0 for m in MODAL_TOKEN_INDEXES}
    new_embeds, new_labels = [], ([] if labels is not None else None)
    masks = {m: [] for m in modal_features}

    for b in range(B):
        cur = ids_np[b]
        cur_labels = labels[b] if labels is not None else None
        modal, start = modal_token_match(cur)
        if modal is None:  # :323-342 the "hacky" no-sentinel path
            new_embeds.append(embed_table[torch.from_numpy(cur.copy()).long()])
            if labels is not None:
                new_labels.append(labels[b])
            for m in modal_features:
                masks[m].append(torch.zeros(len(cur), dtype=mask_dtype))
            continue
        pieces, lab_pieces = [], []
        cur_masks = {m: [] for m in modal_input_keys}
        while modal is not None:
            feats = modal_features[modal][cursor[modal]]
            pieces.append(embed_table[torch.from_numpy(cur[:start].copy()).long()])
            pieces.append(feats)
            for m in cur_masks:
                if m != modal:
                    cur_masks[m].append(torch.zeros(start + len(feats), dtype=mask_dtype))
                else:
                    cur_masks[m].append(torch.zeros(start, dtype=mask_dtype))
                    cur_masks[m].append(torch.ones(len(feats), dtype=mask_dtype))
            if labels is not None:
                lab_pieces.append(cur_labels[:start])
                lab_pieces.append(torch.full((feats.shape[0],), IGNORE_INDEX, dtype=labels.dtype))
                cur_labels = cur_labels[start + 1:]
            cursor[modal] += 1
            cur = cur[start + 1:]
            modal, start = modal_token_match(cur)
        if cur.size > 0:
            pieces.append(embed_table[torch.from_numpy(cur.copy()).long()])
            for m in cur_masks:
                cur_masks[m].append(torch.zeros(len(cur), dtype=mask_dtype))
            if labels is not None:
                lab_pieces.append(cur_labels)
        new_embeds.append(torch.cat(pieces, dim=0))
        for m in cur_masks:
            masks[m].append(torch.cat(cur_masks[m], dim=0))
        if labels is not None:
            new_labels.append(torch.cat(lab_pieces, dim=0))

    ragged = any(x.shape != new_embeds[0].shape for x in new_embeds)
    if ragged:
        if labels is None:
            # :414 ``_new_labels = new_labels`` is only bound under ``if labels is not None``
            raise UnboundLocalError("ragged batch with labels=None (reference multimodal_arch.py:414-429)")
        max_len = max(x.shape[0] for x in new_embeds)
        embeds = torch.stack([torch.cat((x, torch.zeros((max_len - x.shape[0], H), dtype=x.dtype)), 0) for x in new_embeds])
        out_masks = {}
        for m in masks:
            out_masks[m] = torch.stack([torch.cat((x, torch.zeros(max_len - x.shape[0], dtype=mask_dtype)), 0)
                                        for x in masks[m]])
        out_labels = torch.stack([torch.cat((x, torch.full((max_len - x.shape[0],), IGNORE_INDEX, dtype=x.dtype)), 0)
                                  for x in new_labels])
        if attention_mask is not None:
            rows = []
            for am, lab in zip(attention_mask, new_labels):
                left = torch.ones(lab.shape[0] - labels.shape[1], dtype=mask_dtype)
                right = torch.zeros(max_len - lab.shape[0], dtype=mask_dtype)
                rows.append(torch.cat((left, am, right), 0))
            attention_mask = torch.stack(rows)
    else:
        embeds = torch.stack(new_embeds)
        out_masks = {m: torch.stack(v) for m, v in masks.items() if len(v)}
        out_labels = torch.stack(new_labels) if labels is not None else None
        if attention_mask is not None:
            left = torch.ones((attention_mask.shape[0], embeds.shape[1] - S), dtype=mask_dtype)
            attention_mask = torch.cat((left, attention_mask), dim=1)
    if len(out_masks):  # :452-453
        out_masks["default"] = torch.stack([out_masks[k] for k in out_masks]).sum(0) == 0
    else:
        out_masks = None
    return attention_mask, embeds, out_labels, out_masks
