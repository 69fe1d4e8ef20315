"""Prompt -> ``input_ids`` with modality sentinels, and the batch hand-off to the device.

Host-side mirrors of the reference's text plumbing that feeds the hot path:
  modelcompose/constants.py:15-31       MODAL_TOKENS / MODAL_TOKEN_INDEXES / MODAL_TOKEN_MAPPING
  modelcompose/mm_utils.py:60-101       split_string_by_list, tokenizer_modal_token
  modelcompose/data/multimodal_dataset.py:141-170  DataCollatorForSupervisedDataset (ids / labels / attention_mask part)
The reference's collator also runs the per-modality *processors* of the frozen encoders (:172-214); those are out of scope
here (``modal_inputs`` carries encoder features), so ``FeatureCollator`` batches feature tensors.
The reference's batched inference raises on ragged batches when ``labels`` is None: ``bucket_by_length``
groups requests whose spliced length is equal so batches stay rectangular, and ``PinnedBatch`` stages a batch in pinned
host memory so the host->device copies are asynchronous.  No arithmetic happens here.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence

import torch

IGNORE_INDEX = -100
MODAL_TOKENS = {"vision": "<image>", "relrep": "<relrep>", "text": "<text>", "audio": "<audio>", "video": "<video>",
                "point": "<point>"}
MODAL_TOKEN_INDEXES = {"vision": -200, "relrep": -201, "text": -202, "audio": -203, "video": -204, "point": -205}
MODAL_TOKEN_MAPPING = {MODAL_TOKENS[k]: MODAL_TOKEN_INDEXES[k] for k in MODAL_TOKENS}


def split_string_by_list(input_string: str, split_list: Sequence[str]):
    """``[(text, separator or None), ...]`` with the semantics of mm_utils.py:64-78.  The reference grows a chunk one
    character at a time and cuts as soon as any separator is contained in it, i.e. at the separator whose match ENDS first
    (ties: list order).  Here that is computed directly: per separator one ``str.find`` from the current position, the
    smallest end offset wins."""
    pieces, pos, n = [], 0, len(input_string)
    while pos < n:
        cut = None  # (end, start, separator)
        for sep in split_list:
            at = input_string.find(sep, pos) if sep else -1
            if at >= 0 and (cut is None or at + len(sep) < cut[0]):
                cut = (at + len(sep), at, sep)
        if cut is None:
            pieces.append((input_string[pos:], None))
            break
        pieces.append((input_string[pos:cut[1]], cut[2]))
        pos = cut[0]
    return pieces


def tokenizer_modal_token(prompt: str, tokenizer, return_tensors: Optional[str] = None):
    """mm_utils.py:81-101: the text between modality placeholders is tokenised piece by piece and each placeholder becomes
    its negative sentinel id; the BOS the tokenizer puts in front of every piece survives once, at the very start."""
    pieces = split_string_by_list(prompt, list(MODAL_TOKEN_MAPPING))
    tokenised = [tokenizer(text).input_ids for text, _ in pieces]
    has_bos = bool(tokenised) and bool(tokenised[0]) and tokenised[0][0] == tokenizer.bos_token_id
    skip = 1 if has_bos else 0
    input_ids: List[int] = [tokenised[0][0]] if has_bos else []
    for toks, (_, sep) in zip(tokenised, pieces):
        input_ids += toks[skip:]
        if sep is not None:
            input_ids.append(MODAL_TOKEN_MAPPING[sep])
    if return_tensors is None:
        return input_ids
    if return_tensors != "pt":
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return torch.tensor(input_ids, dtype=torch.long)


class FeatureCollator:
    """multimodal_dataset.py:141-170 for instances whose ``modal_inputs`` hold encoder FEATURES: pads ``input_ids`` /
    ``labels`` to the longest of the batch (pad id / IGNORE_INDEX, truncated to ``model_max_length``), builds
    ``attention_mask = input_ids != pad``, and concatenates every modality's feature blocks in instance order — the
    order the splice's batch-global cursor consumes them in (multimodal_arch.py:302,:365)."""

    def __init__(self, pad_token_id: int, model_max_length: int = 2048):
        self.pad_token_id, self.model_max_length = int(pad_token_id), int(model_max_length)

    def __call__(self, instances: Sequence[Dict]) -> Dict:
        ids = [torch.as_tensor(i["input_ids"], dtype=torch.long) for i in instances]
        input_ids = torch.nn.utils.rnn.pad_sequence(ids, batch_first=True, padding_value=self.pad_token_id)
        input_ids = input_ids[:, :self.model_max_length]
        batch = {"input_ids": input_ids, "attention_mask": input_ids.ne(self.pad_token_id)}
        if "labels" in instances[0]:
            labels = [torch.as_tensor(i["labels"], dtype=torch.long) for i in instances]
            batch["labels"] = torch.nn.utils.rnn.pad_sequence(labels, batch_first=True,
                                                              padding_value=IGNORE_INDEX)[:, :self.model_max_length]
        if "modal_inputs" in instances[0]:
            per_modal = defaultdict(list)
            for inst in instances:
                for modal, blocks in inst["modal_inputs"].items():
                    per_modal[modal].extend(blocks if isinstance(blocks, (list, tuple)) else [blocks])
            batch["modal_inputs"] = {m: torch.stack(v, dim=0) for m, v in per_modal.items()}
        return batch


def spliced_length(input_ids: Sequence[int], rows_per_block: Dict[str, int]) -> int:
    """Length of a request after the splice: every sentinel is replaced by its modality's block (feature rows + prefix +
    suffix rows, ``rows_per_block[modal]``); unknown sentinels count as one token (the splice will reject them)."""
    index_to_modal = {v: k for k, v in MODAL_TOKEN_INDEXES.items()}
    n = 0
    for t in input_ids:
        t = int(t)
        n += rows_per_block.get(index_to_modal.get(t, ""), 1) if t < 0 else 1
    return n


def bucket_by_length(instances: Sequence[Dict], batch_size: int, rows_per_block: Optional[Dict[str, int]] = None) -> Iterable[List[int]]:
    """Index groups of at most ``batch_size`` requests with EQUAL token count before and after the splice, in first-seen
    order: every batch is rectangular, which is what the reference's inference path (and the kernels' one-shape
    workspaces) need."""
    buckets: Dict[tuple, List[int]] = {}
    for i, inst in enumerate(instances):
        ids = inst["input_ids"]
        key = (len(ids), spliced_length(ids, rows_per_block or {}))
        b = buckets.setdefault(key, [])
        b.append(i)
        if len(b) == batch_size:
            yield list(b)
            b.clear()
    for b in buckets.values():
        if b:
            yield list(b)


class PinnedBatch:
    """Reusable pinned host staging for one batch shape: ``load(batch)`` copies the collated tensors into pinned buffers,
    ``to(device)`` issues non-blocking host->device copies on the current stream and returns device tensors."""

    def __init__(self):
        self._bufs: Dict[str, torch.Tensor] = {}

    def _stage(self, name: str, t: torch.Tensor) -> torch.Tensor:
        buf = self._bufs.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=torch.cuda.is_available())
            self._bufs[name] = buf
        buf.copy_(t)
        return buf

    def load(self, batch: Dict) -> "PinnedBatch":
        self._batch = {k: self._stage(k, v) for k, v in batch.items() if torch.is_tensor(v)}
        self._modal = {m: self._stage("modal." + m, v) for m, v in batch.get("modal_inputs", {}).items()}
        return self

    def to(self, device) -> Dict:
        out = {k: v.to(device, non_blocking=True) for k, v in self._batch.items()}
        if self._modal:
            out["modal_inputs"] = {m: v.to(device, non_blocking=True) for m, v in self._modal.items()}
        return out
