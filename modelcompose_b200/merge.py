"""Checkpoint merge: host logic of the reference CLI + the CUDA N-source merge behind it.

Host side mirrors ``scripts/model_composition/merge_unimodal_modelcompose.py`` of the reference
(same function names, arguments, on-disk outputs and error behaviour).
All tensor arithmetic goes through the C ABI (``mc_merge_*`` in include/modelcompose_b200.h) —
there is no torch/CPU arithmetic path in this module.
"""
from __future__ import annotations

import argparse
import copy
import ctypes as C
import json
import os
from collections import OrderedDict, defaultdict
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi

# reference merge_unimodal_modelcompose.py:15-21 — first matching key wins
MODAL_DICT = {
    "mm_vision_encoder": "vision",
    "mm_vision_tower": "vision",
    "mm_vision2_encoder": "vision2",
    "mm_vision2_tower": "vision2",
    "mm_video_encoder": "video",
    "mm_audio_encoder": "audio",
    "mm_point_encoder": "point",
}

MODES = {"weighted": _cabi.MC_MERGE_WEIGHTED, "sum": _cabi.MC_MERGE_REF_SUM, "mean": _cabi.MC_MERGE_REF_MEAN}
TIES_FUNCS = {"sum": _cabi.MC_TIES_SUM, "mean": _cabi.MC_TIES_MEAN, "max": _cabi.MC_TIES_MAX}


def get_modal_from_config(config: dict) -> str:
    """reference :22-26."""
    for key, modal in MODAL_DICT.items():
        value = config.get(key) if key in config.keys() else None
        if isinstance(value, str) and len(value) > 0:
            return modal
    assert False, "No modality is recognized, please check the config."


# ------------------------------------------------------------------------------------------------ device API
class MergePlan:
    """Pointer/chunk tables for merging N same-shaped tensor lists resident on ONE device.

    ``sources[s][t]`` and ``outputs[t]`` are contiguous CUDA tensors; ``run(weights, mode)`` enqueues one
    kernel launch on the current stream.  The plan keeps references to the tensors it points at."""

    def __init__(self, sources: Sequence[Sequence[torch.Tensor]], outputs: Sequence[torch.Tensor], tuning: int = 0):
        n_src, n_t = len(sources), len(outputs)
        if not 1 <= n_src <= _cabi.MC_MERGE_MAX_SRC:
            raise ValueError(f"need 1..{_cabi.MC_MERGE_MAX_SRC} sources, got {n_src}")
        src_dtype = sources[0][0].dtype if n_t else torch.bfloat16
        dst_dtype = outputs[0].dtype if n_t else torch.bfloat16
        for s in sources:
            if len(s) != n_t:
                raise ValueError("every source must hold the same number of tensors")
        for t in range(n_t):
            o = outputs[t]
            _check_device_tensor(o, dst_dtype, o.numel(), f"outputs[{t}]")
            for k in range(n_src):
                _check_device_tensor(sources[k][t], src_dtype, o.numel(), f"sources[{k}][{t}]")
        self._keep = (list(map(list, sources)), list(outputs))
        self.n_src, self.n_tensors = n_src, n_t
        self._h = C.c_void_p()
        src_ptrs = _cabi.ptr_array([sources[k][t].data_ptr() for k in range(n_src) for t in range(n_t)])
        dst_ptrs = _cabi.ptr_array([o.data_ptr() for o in outputs])
        numel = _cabi.i64_array([o.numel() for o in outputs])
        _cabi.check(_cabi.lib().mc_merge_plan_create(C.byref(self._h), n_t, n_src, src_ptrs, dst_ptrs, numel,
                                                     _cabi.dtype_code(src_dtype), _cabi.dtype_code(dst_dtype), tuning),
                    "mc_merge_plan_create")

    @property
    def algorithmic_bytes(self) -> int:
        return int(_cabi.lib().mc_merge_plan_bytes(self._h))

    def run(self, weights: Optional[Sequence[float]] = None, mode: str = "weighted") -> None:
        if mode == "weighted":
            if weights is None or len(weights) != self.n_src:
                raise ValueError("weighted merge needs one weight per source")
            w = _cabi.f32_array(weights)
        else:
            w = _cabi.f32_array([1.0] * self.n_src)
        _cabi.check(_cabi.lib().mc_merge_plan_run(self._h, w, MODES[mode], _cabi.current_stream_ptr()),
                    "mc_merge_plan_run")
        _cabi.count_launch()

    def close(self) -> None:
        if self._h:
            _cabi.lib().mc_merge_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_device_tensor(t: torch.Tensor, dtype, numel: int, what: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{what} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype or t.numel() != numel or not t.is_contiguous():
        raise ValueError(f"{what}: expected contiguous {dtype} with {numel} elements, got {t.dtype} {tuple(t.shape)}")


def merge_state_dicts_device(state_dicts: Sequence[Dict[str, torch.Tensor]], weights: Sequence[float],
                             out_dtype=None, mode: str = "weighted") -> Dict[str, torch.Tensor]:
    """Merge N device-resident state dicts with identical keys/shapes: ``out[k] = Σ_m w_m · sd_m[k]``
    (one kernel launch for all tensors).  This is the materialised online-merge-reset blend
     when ``state_dicts = [base, ckpt_1, …]`` and
    ``weights = [1-Σw, w_1, …]``."""
    keys = list(state_dicts[0].keys())
    srcs = [[sd[k].contiguous() for k in keys] for sd in state_dicts]
    outs = [torch.empty_like(t, dtype=out_dtype or t.dtype) for t in srcs[0]]
    plan = MergePlan(srcs, outs)
    plan.run(weights, mode)
    torch.cuda.current_stream().synchronize()
    plan.close()
    return dict(zip(keys, outs))


def merge_host_tensors(tensor_lists: Sequence[Sequence[torch.Tensor]], weights: Optional[Sequence[float]] = None,
                       mode: str = "weighted", out_dtype=None, staging_bytes: int = 0) -> List[torch.Tensor]:
    """Merge HOST tensors through the GPU (``mc_merge_host``: H2D, kernel, D2H pipelined).
    ``tensor_lists[s][t]``: CPU tensors; returns new CPU tensors.  Used by the CLI's ``sum``/``mean``."""
    n_src, n_t = len(tensor_lists), len(tensor_lists[0])
    if not torch.cuda.is_available():
        raise _cabi.McError("merging tensors needs a CUDA device (modelcompose_b200 has no CPU fallback)")
    srcs = [[t.contiguous() for t in lst] for lst in tensor_lists]
    src_dtype = srcs[0][0].dtype if n_t else torch.bfloat16
    for si, lst in enumerate(srcs):
        for ti, (a, b) in enumerate(zip(lst, srcs[0])):
            if a.dtype != src_dtype or a.shape != b.shape or a.is_cuda:
                raise ValueError(f"source {si}, tensor {ti}: sources must be CPU tensors of one dtype and shape per launch "
                                 f"(got {a.dtype} {tuple(a.shape)} vs {src_dtype} {tuple(b.shape)})")
    out_dtype = out_dtype or src_dtype
    outs = [torch.empty(t.shape, dtype=out_dtype) for t in srcs[0]]
    w = _cabi.f32_array(weights if weights is not None else [1.0] * n_src)
    _cabi.check(_cabi.lib().mc_merge_host(
        n_t, n_src, _cabi.ptr_array([srcs[k][t].data_ptr() for k in range(n_src) for t in range(n_t)]),
        _cabi.ptr_array([o.data_ptr() for o in outs]), _cabi.i64_array([o.numel() for o in outs]), w, MODES[mode],
        _cabi.dtype_code(src_dtype), _cabi.dtype_code(out_dtype), staging_bytes), "mc_merge_host")
    return outs


# ------------------------------------------------------------------------------------------------ TIES merge
def ties_kth_rank(d: int, K) -> int:
    """reference ties_merging.py:89-96 — ``K >= 1`` is a percentage; the kept elements are those whose magnitude is at
    least the ``k = d - int(d * K)``-th smallest (1-based).  The reference's ``kthvalue(k)`` raises for k == 0."""
    if K >= 1:
        K = K / 100
    k = d - int(d * K)
    if not 1 <= k <= d:
        raise RuntimeError(f"kthvalue(): selected number k out of range for dimension 1 (k={k}, d={d})")
    return k


def _stats_dict(st: "_cabi.TiesStats", n_src: int) -> dict:
    return {"thresholds": [float(st.threshold[i]) for i in range(n_src)], "n_pos": int(st.n_pos), "n_neg": int(st.n_neg),
            "n_zero": int(st.n_zero), "n_ambiguous": int(st.n_ambiguous), "majority": int(st.majority),
            "full_select_ran": bool(st.full_select_ran), "fix_pass_ran": int(st.fix_pass_ran)}  # 0 none, 1 sparse fix-up, 2 dense re-merge


def _metrics_dict(m: "_cabi.InterferenceMetrics", n_src: int) -> dict:
    return {"L2": float(m.l2), "Cosine": float(m.cosine), "SSD": float(m.ssd), "TSSD": float(m.tssd),
            "ssd_elements": int(m.ssd_elements), "tssd_elements": int(m.tssd_elements),
            "thresholds": [float(m.threshold[i]) for i in range(n_src)]}


class TiesPlan:
    """TIES merge (trim / elect sign / disjoint merge, reference ties_merging.py:161-179) of N same-shaped tensor lists
    resident on ONE device.  ``outputs`` must be float32 for ``func='mean'`` (the reference's result dtype) and the
    source dtype otherwise.  ``run(K, func)`` enqueues every pass on the current stream; nothing synchronises."""

    def __init__(self, sources: Sequence[Sequence[torch.Tensor]], outputs: Optional[Sequence[torch.Tensor]] = None):
        n_src, n_t = len(sources), len(sources[0])
        if not 1 <= n_src <= _cabi.MC_MERGE_MAX_SRC:
            raise ValueError(f"need 1..{_cabi.MC_MERGE_MAX_SRC} sources, got {n_src}")
        if n_t == 0:
            raise ValueError("nothing to merge")
        src_dtype = sources[0][0].dtype
        dst_dtype = outputs[0].dtype if outputs is not None else src_dtype
        for s in sources:
            if len(s) != n_t:
                raise ValueError("every source must hold the same number of tensors")
        if outputs is not None and len(outputs) != n_t:
            raise ValueError("one output per tensor")
        for t in range(n_t):
            numel = sources[0][t].numel()
            if outputs is not None:
                _check_device_tensor(outputs[t], dst_dtype, numel, f"outputs[{t}]")
            for k in range(n_src):
                _check_device_tensor(sources[k][t], src_dtype, numel, f"sources[{k}][{t}]")
        self._keep = (list(map(list, sources)), list(outputs) if outputs is not None else None)
        self.n_src, self.n_tensors, self.src_dtype = n_src, n_t, src_dtype
        self._h = C.c_void_p()
        _cabi.check(_cabi.lib().mc_ties_plan_create(
            C.byref(self._h), n_t, n_src, _cabi.ptr_array([sources[k][t].data_ptr() for k in range(n_src) for t in range(n_t)]),
            _cabi.ptr_array([o.data_ptr() for o in outputs]) if outputs is not None else None,
            _cabi.i64_array([t.numel() for t in sources[0]]),
            _cabi.dtype_code(src_dtype), _cabi.dtype_code(dst_dtype)), "mc_ties_plan_create")
        self.elements = int(_cabi.lib().mc_ties_plan_elements(self._h))

    @property
    def algorithmic_bytes(self) -> int:
        return int(_cabi.lib().mc_ties_plan_bytes(self._h))

    def run(self, K=20, func: str = "mean") -> None:
        _cabi.check(_cabi.lib().mc_ties_plan_run(self._h, ties_kth_rank(self.elements, K), TIES_FUNCS[func],
                                                 _cabi.current_stream_ptr()), "mc_ties_plan_run")
        # 16-bit sources: sample + bracket, count + select, full-range pass (exits at once), merge, fix-up, re-merge; fp32: init and three
        # full-range passes instead of the first three
        _cabi.count_launch(6 if self.src_dtype != torch.float32 else 7)

    def metrics(self, reset_thresh=50) -> dict:
        """Parameter-interference metrics of the plan's sources (reference calculate_metrics.py:26-37,53-64):
        L2 / cosine distance of the first two sources, soft sign dissimilarity before and after the top-``reset_thresh``
        trim.  Synchronises the current stream."""
        m = _cabi.InterferenceMetrics()
        _cabi.check(_cabi.lib().mc_ties_plan_metrics(self._h, ties_kth_rank(self.elements, reset_thresh), C.byref(m),
                                                     _cabi.current_stream_ptr()), "mc_ties_plan_metrics")
        _cabi.count_launch(8)
        return _metrics_dict(m, self.n_src)

    def stats(self) -> dict:
        st = _cabi.TiesStats()
        _cabi.check(_cabi.lib().mc_ties_plan_stats(self._h, C.byref(st), _cabi.current_stream_ptr()), "mc_ties_plan_stats")
        return _stats_dict(st, self.n_src)

    def close(self) -> None:
        if self._h:
            _cabi.lib().mc_ties_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _promoted_sources(tensor_lists: Sequence[Sequence[torch.Tensor]], names: Optional[Sequence[str]] = None):
    """Contiguous CPU copies of ``tensor_lists[s][t]`` in ONE dtype: the reference flattens every checkpoint with
    ``torch.cat`` and stacks the vectors with ``vstack`` (ties_merging.py:12-19,:188-190; calculate_metrics casts with
    ``.float()``), both of which type-promote, so a float32 projector beside bf16 adapters makes the whole problem float32."""
    n_t = len(tensor_lists[0])
    common = None
    for si, lst in enumerate(tensor_lists):
        if len(lst) != n_t:
            raise ValueError("every source must hold the same number of tensors")
        for ti, (a, b) in enumerate(zip(lst, tensor_lists[0])):
            what = names[ti] if names is not None else f"tensor {ti}"
            if a.is_cuda or a.shape != b.shape:
                raise ValueError(f"source {si}, {what}: sources must be CPU tensors of identical shape per key "
                                 f"(got {tuple(a.shape)} vs {tuple(b.shape)})")
            if not a.dtype.is_floating_point:
                raise ValueError(f"source {si}, {what}: {a.dtype} is not a floating-point dtype")
            common = a.dtype if common is None else torch.promote_types(common, a.dtype)
    if common not in (torch.float32, torch.float16, torch.bfloat16):
        raise ValueError(f"sources promote to {common}; the kernels take float32, float16 or bfloat16")
    return [[t.to(common).contiguous() for t in lst] for lst in tensor_lists], common


def ties_merge_host_tensors(tensor_lists: Sequence[Sequence[torch.Tensor]], K=20, func: str = "mean",
                            names: Optional[Sequence[str]] = None, outputs: Optional[Sequence[torch.Tensor]] = None):
    """TIES-merge HOST tensors through the GPU (``mc_ties_host``).  ``tensor_lists[s][t]``: CPU tensors, same shapes across
    sources; mixed dtypes are promoted to one (as the reference's flatten does).  Returns (CPU tensors, stats dict);
    float32 outputs for ``mean``.  ``outputs``: write into these CPU tensors (e.g. pinned ones) instead of new ones."""
    if not torch.cuda.is_available():
        raise _cabi.McError("merging tensors needs a CUDA device (modelcompose_b200 has no CPU fallback)")
    n_src, n_t = len(tensor_lists), len(tensor_lists[0])
    srcs, src_dtype = _promoted_sources(tensor_lists, names)
    out_dtype = torch.float32 if func == "mean" else src_dtype
    if outputs is None:
        outs = [torch.empty(t.shape, dtype=out_dtype) for t in srcs[0]]
    else:
        outs = list(outputs)
        if len(outs) != n_t:
            raise ValueError("one output per tensor")
        for ti, (o, t) in enumerate(zip(outs, srcs[0])):
            if o.is_cuda or o.dtype != out_dtype or o.shape != t.shape or not o.is_contiguous():
                raise ValueError(f"outputs[{ti}]: need a contiguous CPU tensor of {out_dtype} {tuple(t.shape)}, "
                                 f"got {o.dtype} {tuple(o.shape)}")
    d = sum(o.numel() for o in outs)
    st = _cabi.TiesStats()
    _cabi.check(_cabi.lib().mc_ties_host(
        n_t, n_src, _cabi.ptr_array([srcs[k][t].data_ptr() for k in range(n_src) for t in range(n_t)]),
        _cabi.ptr_array([o.data_ptr() for o in outs]), _cabi.i64_array([o.numel() for o in outs]), ties_kth_rank(d, K),
        TIES_FUNCS[func], _cabi.dtype_code(src_dtype), C.byref(st)), "mc_ties_host")
    return outs, _stats_dict(st, n_src)


def interference_metrics_host(tensor_lists: Sequence[Sequence[torch.Tensor]], reset_thresh=50) -> dict:
    """Interference metrics of HOST tensors through the GPU (``mc_interference_host``); ``tensor_lists[s][t]`` as
    ``ties_merge_host_tensors``.  The reference indexes the second source unconditionally: fewer than two raise IndexError."""
    if not torch.cuda.is_available():
        raise _cabi.McError("interference metrics need a CUDA device (modelcompose_b200 has no CPU fallback)")
    n_src, n_t = len(tensor_lists), len(tensor_lists[0])
    if n_src < 2:
        raise IndexError("index 1 is out of bounds for dimension 0 with size 1")
    srcs, src_dtype = _promoted_sources(tensor_lists)
    d = sum(t.numel() for t in srcs[0])
    m = _cabi.InterferenceMetrics()
    _cabi.check(_cabi.lib().mc_interference_host(
        n_t, n_src, _cabi.ptr_array([srcs[k][t].data_ptr() for k in range(n_src) for t in range(n_t)]),
        _cabi.i64_array([t.numel() for t in srcs[0]]), ties_kth_rank(d, reset_thresh), _cabi.dtype_code(src_dtype), C.byref(m)),
        "mc_interference_host")
    return _metrics_dict(m, n_src)


def convert_delta_to_ft(delta_weights: Dict[str, List[torch.Tensor]]):
    """``{key: [tensor of every checkpoint that has the key]}`` -> (one state dict per checkpoint holding the keys ALL of them
    share, dict of the keys a single checkpoint contributes), the split the reference makes before flattening
    (ties_merging.py:224-250).  Any other multiplicity is a caller error, as there."""
    n_ckpt = max((len(tensors) for tensors in delta_weights.values()), default=0)
    assert n_ckpt > 0
    shared = [dict() for _ in range(n_ckpt)]
    single = {}
    for key, tensors in delta_weights.items():
        if len(tensors) == n_ckpt:
            for state, t in zip(shared, tensors):
                state[key] = t
        else:
            assert len(tensors) == 1
            single[key] = tensors[0]
    return shared, single


def do_merging(ft_checks: Sequence[Dict[str, torch.Tensor]], K=20, merge_func: str = "dis-mean", lamda=1):
    """reference ties_merging.py:182-222 on the GPU: same inputs (a list of state dicts with identical keys), same output
    (an ordered dict in sorted-key order, tensors shaped like ``ft_checks[0]``; float32 for dis-mean).  The reference
    flattens each source into one vector; the kernels work on the tensors in place — the trim threshold and the
    majority sign are statistics over all of them, everything else is elementwise."""
    if lamda != 1:
        raise NotImplementedError("the reference only ever calls do_merging with lamda = 1")
    keys = sorted(ft_checks[0])
    for check in ft_checks[1:]:
        if set(check.keys()) != set(keys):
            raise ValueError("Differing parameter names in models. "
                             f"The different parameters are {set(keys).symmetric_difference(set(check.keys()))}")
    func = merge_func.split("-")[-1]
    if func not in TIES_FUNCS:
        raise ValueError(f"Merge method {func} is not defined.")
    outs, _ = ties_merge_host_tensors([[check[k] for k in keys] for check in ft_checks], K=K, func=func, names=keys)
    return OrderedDict(zip(keys, outs))


# ------------------------------------------------------------------------------------------------ CLI host logic
def _load_checkpoint_dir(filepath: str):
    """reference :31-36 — adapter_model.bin (fallback mm_projector.bin) + config.json."""
    adapter_path = os.path.join(filepath, "adapter_model.bin")
    if not os.path.exists(adapter_path):
        adapter_path = os.path.join(filepath, "mm_projector.bin")
    weights = torch.load(adapter_path, map_location=torch.device("cpu"))
    with open(os.path.join(filepath, "config.json")) as f:
        config = json.load(f)
    return weights, config


def _elementwise_strategy(weights_to_merge: Dict[str, List[torch.Tensor]], strategy: str) -> Dict[str, torch.Tensor]:
    """reference :105-112 (`sum`, `mean`) on the GPU.  Keys are grouped by source count so each group is one
    multi-tensor launch; results are bit-identical to the reference's per-add storage-dtype rounding."""
    merged: Dict[str, torch.Tensor] = {}
    by_count: Dict[tuple, List[str]] = defaultdict(list)
    promoted: Dict[str, List[torch.Tensor]] = {}
    for key, tensors in weights_to_merge.items():
        dtypes = {t.dtype for t in tensors}
        if len(dtypes) > 1:
            # Python's sum() adds left to right and torch promotes each add.  With two sources the only add of two non-zero
            # operands runs in the promoted dtype, so casting both up first gives the same bits; with more, earlier partial
            # sums would round in the narrower dtype — not reproduced here, so say so instead of returning other bits.
            if len(tensors) > 2:
                raise ValueError(f"{key}: {len(tensors)} sources in mixed dtypes {sorted(map(str, dtypes))} — cast the "
                                 "checkpoints to one dtype before merging")
            common = torch.promote_types(tensors[0].dtype, tensors[1].dtype)
            tensors = [t.to(common) for t in tensors]
        promoted[key] = tensors
        by_count[(len(tensors), tensors[0].dtype)].append(key)
    for (n_src, _), keys in by_count.items():
        lists = [[promoted[k][s] for k in keys] for s in range(n_src)]
        outs = merge_host_tensors(lists, mode=strategy)
        merged.update(zip(keys, outs))
    return {k: merged[k] for k in weights_to_merge}  # first-seen key order, as the reference dict has


def merge_checkpoints(filepaths, output_path, strategy="sum", K=20):
    """Drop-in for reference ``merge_checkpoints`` (:28-145): same inputs, same three output files.

    Strategies: ``online-merge-[reset-]…`` (key rename + coefficient string into config.json; no arithmetic,
    exactly as the reference), ``sum`` / ``mean`` (N-source elementwise merge on the GPU), ``ties-{sum,mean,max}``
    (TIES merge on the GPU, ``-K`` = percentage of entries kept), and the ``convert-`` prefix (``same``-strategy
    checkpoints re-keyed per modality) with its ``convert-drop-*`` variant."""
    configs, weights_to_merge = [], defaultdict(list)
    for filepath in filepaths:
        adapter_weights, modal_config = _load_checkpoint_dir(filepath)
        configs.append(modal_config)
        for key in adapter_weights:
            weights_to_merge[key].append(adapter_weights[key])

    merged_weights = None
    if strategy.startswith("convert-"):  # :42-71 — 'same'-strategy checkpoints to 'modal+language'
        strategy = strategy.replace("convert-", "")
        for config in configs:
            if "lora_strategy" in config:
                assert config["lora_strategy"] == "same"
                config["lora_strategy"] = "modal+language"
        modal_types = [get_modal_from_config(config) for config in configs]
        convert_weights_to_merge = defaultdict(list)
        for key in weights_to_merge:
            if ".default" in key:
                for i in range(len(modal_types)):
                    convert_weights_to_merge[key.replace("default", modal_types[i])].append(copy.deepcopy(weights_to_merge[key][i]))
        if strategy.startswith("drop-"):
            ft_checks, uniques = convert_delta_to_ft(weights_to_merge)
            merged_weights = do_merging(ft_checks, K=K, merge_func=strategy.replace("drop-", "dis-"))
            merged_weights.update(uniques)
            for k in convert_weights_to_merge:
                convert_weights_to_merge[k] = convert_weights_to_merge[k][0]
            merged_weights.update(convert_weights_to_merge)
        else:
            weights_to_merge.update(convert_weights_to_merge)
    print(strategy, strategy.startswith("ties-"))

    if strategy.startswith("ties-"):  # :75-93
        assert strategy.replace("ties-", "") in ["sum", "mean", "max"]
        ft_checks, uniques = convert_delta_to_ft(weights_to_merge)
        merge_func = strategy.replace("ties-", "dis-")
        merged_weights = do_merging(ft_checks, K=K, merge_func=merge_func)
        merged_weights.update(uniques)
        strategy = f"{merge_func}-{K}"
        assert sorted(weights_to_merge) == sorted(merged_weights), "the keys should be the same"
    elif strategy.startswith("online-merge-"):
        merged_weights = {}
        modal_names = [get_modal_from_config(config) for config in configs]
        for key, tensors in weights_to_merge.items():
            if len(tensors) == 1:
                merged_weights[key] = tensors[0]
                continue
            assert "default" in key
            for modal_name, weight in zip(modal_names, tensors):
                merged_weights[key.replace("default", f"default-{modal_name}")] = weight
    elif strategy in ("sum", "mean"):
        merged_weights = _elementwise_strategy(weights_to_merge, strategy)
    else:
        print(f"Merge strategy [{strategy}] not implemented, DO NOTHING.")
        if merged_weights is None:
            # the reference falls through to torch.save(merged_weights) with the name unbound (:113-115,:139); after a
            # `convert-drop-*` it is bound and the run completes
            raise UnboundLocalError("cannot access local variable 'merged_weights' where it is not associated with a value")

    merged_configs = {}
    for config in configs:
        for key, value in config.items():
            merged_configs[key] = (merged_configs[key] or value) if key in merged_configs else value
        if strategy.startswith("online-merge-"):  # consumed on the first config only (:124-129)
            strategy = strategy.replace("online-merge-", "")
            if strategy.startswith("reset-"):
                merged_configs["reset_scaling_weights"] = strategy.replace("reset-", "")
            else:
                merged_configs["merge_default_weights"] = strategy
    for config in configs:
        modal_name = get_modal_from_config(config)
        merged_configs[f"{modal_name}_lora_alpha"] = config["lora_alpha"]
        merged_configs[f"{modal_name}_lora_r"] = config["lora_r"]

    os.makedirs(output_path, exist_ok=True)
    torch.save(merged_weights, os.path.join(output_path, "adapter_model.bin"))
    with open(os.path.join(output_path, "config.json"), "w") as f:
        json.dump(merged_configs, f, indent=4)
    with open(os.path.join(output_path, "merge_info.txt"), "w") as fout:
        inputs = "\n".join(filepaths)
        fout.write(f"Inputs:\n{inputs}\n\nOutput({strategy}):{output_path}")
    print(f"Merged checkpoints saved to {output_path}")
    return merged_weights, merged_configs


def write_dense_text_weights(merged_weights: Dict[str, torch.Tensor], merged_configs: dict, model_base: str, output_path: str,
                             device="cuda", dtype=None) -> Dict[str, torch.Tensor]:
    """The online-merge-reset composition as DENSE weights of the language model's text path, written next to the adapter
    checkpoint as ``effective_weights.bin``: for every decoder linear
        W_eff = (1 − Σ_m w_m) · W_base + Σ_m w_m · (W_base + s · B_{default-m} A_{default-m}),
    the weighted combination of the N modality checkpoints with the reset coefficients w_m of the strategy string — the
    materialisation the reference's tooling performs on the CPU (scripts/convert_to_multimodal.py:111-113,
    scripts/model_composition/delta_weights_compare.py:24-31,61) and the blend its model evaluates per forward
    (multimodal_llama.py:130-149).  Runs on the GPU: rank-r GEMMs for the dense unimodal checkpoints, the N-source merge kernel
    for the blend (``materialize.effective_weights``).  Keys follow the base checkpoint (``model.layers.{l}.….weight``)."""
    from . import builder as BD
    from . import materialize as MZ
    from . import model as MD
    if not torch.cuda.is_available():
        raise _cabi.McError("materialising the merged weights needs a CUDA device (modelcompose_b200 has no CPU fallback)")
    cfg = MD.MultimodalConfig.from_dict(merged_configs)
    modal_names = MD.infer_modals(cfg)
    names, scaling, default_names = MD.adapter_scaling(modal_names, cfg.lora_r, cfg.lora_alpha, cfg.reset_scaling_weights)
    if default_names is None:
        raise ValueError("nothing to materialise: the strategy carries no default-<modal> coefficients "
                         "(use --strategy online-merge-reset-default-<modal>=w,...)")
    base = BD._load_base_state_dict(model_base)
    out: Dict[str, torch.Tensor] = {}
    for key, W in base.items():
        if not key.endswith(".weight") or W.dim() != 2 or ".layers." not in key:
            continue
        stem = key[:-len(".weight")]
        A = {a: merged_weights[f"{stem}.lora_A.{a}.weight"] for a in default_names if f"{stem}.lora_A.{a}.weight" in merged_weights}
        if not A:
            continue
        dt = dtype or W.dtype
        if dt not in (torch.float16, torch.bfloat16):
            dt = torch.float16   # the reference loads the language model in fp16 (builder.py:185)
        Wd = W.to(device=device, dtype=dt).contiguous()
        Ad = {a: t.to(device=device, dtype=dt).contiguous() for a, t in A.items()}
        Bd = {a: merged_weights[f"{stem}.lora_B.{a}.weight"].to(device=device, dtype=dt).contiguous() for a in A}
        (Weff,) = MZ.effective_weights(Wd, Ad, Bd, scaling, ["default"], default_names, cfg.lora_alpha / cfg.lora_r)
        out[key] = Weff.cpu()
    torch.cuda.synchronize()
    torch.save(out, os.path.join(output_path, "effective_weights.bin"))
    print(f"Dense text-path weights of {len(out)} linears saved to {os.path.join(output_path, 'effective_weights.bin')}")
    return out


def main(argv=None):
    """reference :151-159 — identical flags, plus ``--materialize-base DIR``: after an ``online-merge-reset-…`` merge also
    write the composed text-path weights as a dense checkpoint (``write_dense_text_weights``)."""
    parser = argparse.ArgumentParser(description="Merge multiple torch checkpoints")
    parser.add_argument("filepaths", nargs="+", help="List of checkpoint file paths to merge")
    parser.add_argument("-o", "--output", default="merged_checkpoint.pth", help="Output file path")
    parser.add_argument("--strategy", default="sum", help="Merge strategy")
    parser.add_argument("-K", default=20, type=int, help="K for ties-merging")
    parser.add_argument("--materialize-base", default=None, metavar="DIR",
                        help="base language model directory: also write the merged (reset-blended) dense weights of the text path")
    args = parser.parse_args(argv)
    merged_weights, merged_configs = merge_checkpoints(args.filepaths, args.output, args.strategy, args.K)
    if args.materialize_base:
        write_dense_text_weights(merged_weights, merged_configs, args.materialize_base, args.output)


if __name__ == "__main__":
    main()
