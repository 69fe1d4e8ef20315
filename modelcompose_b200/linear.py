"""Grouped / modality-routed linears: Python handles over ``mc_linear_*`` (include/modelcompose_b200.h).

The arithmetic lives in csrc/mc_linear.cu (tcgen05 + TMA); this module only marshals pointers and builds the
packed adapter layouts the kernel consumes.  Reference semantics: ``LocalLoraLinear.forward``
(modelcompose/model/language_model/multimodal_llama.py:120-160) routed per token by the modality masks
(:262-268, :380-390), and the ``mlp2x_gelu`` / ``linear`` projectors (multimodal_projector/builder.py:202-219).
There is no torch fallback: every product goes through the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_ROWMASK, EPI_RESIDUAL, EPI_SILU_MUL, EPI_ROPE = 0, 1, 2, 3, 4, 5, 6
MAX_PROBLEMS = 4
MAX_SEGMENTS = 8
TILE_M = 128
K_BLOCK = 64


_PROFILE = None  # list of (start, end) CUDA events around every linear launch while profiling is on
# development switch: 1 = row-per-thread epilogues everywhere (tuning bit 17), the round-1 form, for A/B runs against the
# coalesced shared-memory-transposed NONE / RESIDUAL / SILU_MUL epilogues
import os as _os
EPI_ROWWISE = _os.environ.get("MC_LINEAR_EPI_ROWWISE", "0") != "0"


def start_profile() -> None:
    """Record a CUDA-event pair (on the launching stream) around every linear launch until ``stop_profile``."""
    global _PROFILE
    _PROFILE = []


def stop_profile():
    """Returns (total milliseconds inside linear launches, number of launches); synchronises."""
    global _PROFILE
    ev, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev), len(ev)


class LinearDesc(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K0", C.c_int32), ("K1", C.c_int32),
                ("A0", C.c_void_p), ("lda0", C.c_int64), ("B0", C.c_void_p), ("ldb0", C.c_int64),
                ("A1", C.c_void_p), ("lda1", C.c_int64), ("B1", C.c_void_p), ("ldb1", C.c_int64),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("ldr", C.c_int64), ("col_scale", C.c_void_p), ("row_group", C.c_void_p),
                ("mtile_mask", C.c_void_p), ("group_cols", C.POINTER(C.c_int32)), ("n_groups", C.c_int32),
                ("epilogue", C.c_int32), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("rope_pos", C.c_void_p),
                ("rope_seq_len", C.c_int32), ("rope_head_dim", C.c_int32), ("c_rowmap", C.c_void_p),
                ("seg_start", C.c_void_p), ("n_seg", C.c_int32), ("B0_seg", C.POINTER(C.c_void_p))]


def _mat(t: torch.Tensor, what: str, dtype=None):
    if not t.is_cuda:
        raise ValueError(f"{what} must be a CUDA tensor (modelcompose_b200 has no CPU fallback)")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{what} must be 2-D with unit stride along the last dimension, got {tuple(t.shape)} / {t.stride()}")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{what}: expected {dtype}, got {t.dtype}")
    return t


@dataclass
class Problem:
    """One ``C = epilogue(A0·B0ᵀ + A1·B1ᵀ)`` of a launch; tensors are 2-D row-major views (last stride 1)."""
    A0: torch.Tensor
    B0: torch.Tensor
    C: torch.Tensor
    A1: Optional[torch.Tensor] = None
    B1: Optional[torch.Tensor] = None
    bias: Optional[torch.Tensor] = None
    residual: Optional[torch.Tensor] = None
    col_scale: Optional[torch.Tensor] = None   # fp32 [N]
    row_group: Optional[torch.Tensor] = None   # uint8 [M]
    mtile_mask: Optional[torch.Tensor] = None  # int32 [ceil(M/128)]
    group_cols: Optional[Sequence[int]] = None
    epilogue: int = EPI_NONE
    rope: Optional[tuple] = None               # EPI_ROPE: (cos table, sin table, int32 position scalar or None, seq_len, head_dim)
    c_rowmap: Optional[torch.Tensor] = None    # int32 [M]: problem row m is written to row c_rowmap[m] of C
    # segmented problem (grouped GEMM over materialised per-group weights): rows [seg_start[g], seg_start[g+1]) use B0_groups[g]
    seg_start: Optional[torch.Tensor] = None   # CUDA int32 [len(B0_groups) + 1], rewritten by the caller per batch
    B0_groups: Optional[Sequence[torch.Tensor]] = None


class LinearPlan:
    """TMA descriptors + tile schedule for up to 4 problems run as one launch.  Keeps the tensors alive."""

    def __init__(self, problems: Sequence[Problem], tuning: int = 0):
        if not 1 <= len(problems) <= MAX_PROBLEMS:
            raise ValueError(f"1..{MAX_PROBLEMS} problems per launch")
        dtype = problems[0].A0.dtype
        descs = (LinearDesc * len(problems))()
        self._keep = []
        for i, p in enumerate(problems):
            A0, B0, Cm = _mat(p.A0, "A0", dtype), _mat(p.B0, "B0", dtype), _mat(p.C, "C", dtype)
            M, K0 = A0.shape
            N = B0.shape[0]
            if B0.shape[1] != K0 or tuple(Cm.shape) != (M, N):
                raise ValueError(f"problem {i}: shapes A0 {tuple(A0.shape)} B0 {tuple(B0.shape)} C {tuple(Cm.shape)} do not match")
            d = descs[i]
            d.M, d.N, d.K0, d.K1 = M, N, K0, 0
            d.A0, d.lda0, d.B0, d.ldb0 = A0.data_ptr(), A0.stride(0), B0.data_ptr(), B0.stride(0)
            d.C, d.ldc = Cm.data_ptr(), Cm.stride(0)
            if p.A1 is not None:
                A1, B1 = _mat(p.A1, "A1", dtype), _mat(p.B1, "B1", dtype)
                if A1.shape[0] != M or B1.shape[0] != N or A1.shape[1] != B1.shape[1]:
                    raise ValueError(f"problem {i}: A1 {tuple(A1.shape)} / B1 {tuple(B1.shape)} do not match M={M}, N={N}")
                d.K1, d.A1, d.lda1, d.B1, d.ldb1 = A1.shape[1], A1.data_ptr(), A1.stride(0), B1.data_ptr(), B1.stride(0)
            if p.bias is not None:
                if p.bias.dtype != dtype or p.bias.numel() != N or not p.bias.is_cuda:
                    raise ValueError(f"problem {i}: bias must be CUDA {dtype} with {N} elements")
                d.bias = p.bias.data_ptr()
            if p.residual is not None:
                R = _mat(p.residual, "residual", dtype)
                d.residual, d.ldr = R.data_ptr(), R.stride(0)
            if p.col_scale is not None:
                if p.col_scale.dtype != torch.float32 or p.col_scale.numel() != N:
                    raise ValueError(f"problem {i}: col_scale must be fp32 [N]")
                d.col_scale = p.col_scale.data_ptr()
            if p.row_group is not None:
                if p.row_group.dtype != torch.uint8 or p.row_group.numel() != M:
                    raise ValueError(f"problem {i}: row_group must be uint8 [M]")
                d.row_group = p.row_group.data_ptr()
            if p.mtile_mask is not None:
                if p.mtile_mask.dtype != torch.int32 or p.mtile_mask.numel() != (M + TILE_M - 1) // TILE_M:
                    raise ValueError(f"problem {i}: mtile_mask must be int32 [ceil(M/128)]")
                d.mtile_mask = p.mtile_mask.data_ptr()
            if p.group_cols is not None:
                arr = (C.c_int32 * len(p.group_cols))(*[int(x) for x in p.group_cols])
                self._keep.append(arr)
                d.group_cols, d.n_groups = arr, len(p.group_cols) - 1
            d.epilogue = p.epilogue
            if p.rope is not None:
                cos, sin, pos, seq_len, head_dim = p.rope
                if cos.dtype != dtype or sin.dtype != dtype or not cos.is_cuda or cos.shape[-1] != head_dim:
                    raise ValueError(f"problem {i}: rope tables must be CUDA {dtype} [positions, head_dim]")
                d.rope_cos, d.rope_sin = cos.data_ptr(), sin.data_ptr()
                d.rope_pos = None if pos is None else pos.data_ptr()
                d.rope_seq_len, d.rope_head_dim = int(seq_len), int(head_dim)
            if p.B0_groups is not None:
                G = len(p.B0_groups)
                if not 1 <= G <= MAX_SEGMENTS or p.seg_start is None or p.seg_start.dtype != torch.int32 or not p.seg_start.is_cuda \
                        or p.seg_start.numel() != G + 1 or not p.seg_start.is_contiguous() or p.A1 is not None:
                    raise ValueError(f"problem {i}: a segmented problem takes 1..{MAX_SEGMENTS} B0_groups, a contiguous CUDA int32 "
                                     "seg_start of len(B0_groups) + 1 entries, and no A1 / B1")
                for g, Bg in enumerate(p.B0_groups):
                    Bg = _mat(Bg, f"B0_groups[{g}]", dtype)
                    if tuple(Bg.shape) != tuple(B0.shape) or Bg.stride(0) != B0.stride(0):
                        raise ValueError(f"problem {i}: B0_groups[{g}] must match B0 in shape and row stride")
                arr = (C.c_void_p * G)(*[Bg.data_ptr() for Bg in p.B0_groups])
                self._keep.append(arr)
                d.seg_start, d.n_seg, d.B0_seg = p.seg_start.data_ptr(), G, arr
            if p.c_rowmap is not None:
                if p.c_rowmap.dtype != torch.int32 or p.c_rowmap.numel() != M or not p.c_rowmap.is_cuda or not p.c_rowmap.is_contiguous():
                    raise ValueError(f"problem {i}: c_rowmap must be a contiguous CUDA int32 [M]")
                d.c_rowmap = p.c_rowmap.data_ptr()
            self._keep.append(p)
        self._h = C.c_void_p()
        _cabi.check(_cabi.lib().mc_linear_plan_create(C.byref(self._h), descs, len(problems), _cabi.dtype_code(dtype),
                                                      int(tuning) | ((1 << 17) if EPI_ROWWISE else 0)), "mc_linear_plan_create")

    @property
    def flops(self) -> float:
        return float(_cabi.lib().mc_linear_plan_flops(self._h))

    def run(self) -> None:
        if _PROFILE is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        _cabi.check(_cabi.lib().mc_linear_plan_run(self._h, _cabi.current_stream_ptr()), "mc_linear_plan_run")
        _cabi.count_launch()
        if _PROFILE is not None:
            b.record()
            _PROFILE.append((a, b))

    def close(self) -> None:
        if getattr(self, "_h", None):
            _cabi.lib().mc_linear_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def route_tile_masks(row_group: torch.Tensor, out: Optional[torch.Tensor] = None, coarsen: int = 1) -> torch.Tensor:
    """int32 [ceil(M/128)]: bit g set iff some row of the 128-row tile belongs to group g.  ``coarsen`` = 2 / 4 gives every
    tile of an aligned run of 2 / 4 tiles the union of the run: the CTA-pair kernels skip LoRA k-blocks per 256 / 512 rows,
    so the down-projection must have produced (zeros in) the rank columns of every group present in that many rows."""
    if not row_group.is_cuda or row_group.dtype != torch.uint8 or not row_group.is_contiguous():
        raise ValueError("row_group must be a contiguous CUDA uint8 tensor")
    M = row_group.numel()
    n = (M + TILE_M - 1) // TILE_M
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=row_group.device)
    if coarsen not in (1, 2, 4):
        raise ValueError("coarsen must be 1, 2 or 4")
    _cabi.check(_cabi.lib().mc_route_tile_masks_coarse(row_group.data_ptr(), M, out.data_ptr(), int(coarsen), _cabi.current_stream_ptr()),
                "mc_route_tile_masks")
    _cabi.count_launch()
    return out


def route_permutation(modal_id: torch.Tensor, lut: Optional[Sequence[int]], n_groups: int, perm: torch.Tensor, inv_perm: torch.Tensor,
                      row_group: torch.Tensor, seg_start: torch.Tensor, group_seq: Optional[torch.Tensor] = None) -> None:
    """Stable counting sort of the batch's rows by routing group (``mc_route_permutation``): fills ``perm`` / ``inv_perm`` (int32
    [T]), ``row_group`` (uint8 [T], buffer order), ``seg_start`` (int32 [n_groups + 1]) and optionally ``group_seq`` (uint8 [T],
    the routing group of every sequence-order row).  ``lut`` maps the splice's modality ids to routing groups (None = identity)."""
    T = modal_id.numel()
    if modal_id.dtype != torch.uint8 or not modal_id.is_cuda or not modal_id.is_contiguous():
        raise ValueError("modal_id must be a contiguous CUDA uint8 tensor")
    for t, dt, n, what in ((perm, torch.int32, T, "perm"), (inv_perm, torch.int32, T, "inv_perm"), (row_group, torch.uint8, T, "row_group"),
                           (seg_start, torch.int32, n_groups + 1, "seg_start")):
        if t.dtype != dt or t.numel() != n or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{what} must be a contiguous CUDA {dt} tensor with {n} elements")
    if group_seq is not None and (group_seq.dtype != torch.uint8 or group_seq.numel() != T or not group_seq.is_contiguous()):
        raise ValueError("group_seq must be a contiguous CUDA uint8 tensor with one entry per row")
    lut_b = None if lut is None else bytes(int(x) for x in lut)
    _cabi.check(_cabi.lib().mc_route_permutation(modal_id.data_ptr(), T, lut_b, 0 if lut is None else len(lut_b), int(n_groups),
                                                 perm.data_ptr(), inv_perm.data_ptr(), row_group.data_ptr(), seg_start.data_ptr(),
                                                 None if group_seq is None else group_seq.data_ptr(), _cabi.current_stream_ptr()),
                "mc_route_permutation")
    _cabi.count_launch()


def attention_causal(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, batch: int, seq_len: int,
                     n_heads: int, softmax_scale: float, out_rowmap: Optional[torch.Tensor] = None, tuning: int = 0) -> torch.Tensor:
    """Causal self-attention over ``[batch * seq_len, n_heads * 128]`` projection outputs (RoPE already applied); the output
    of token t lands in row ``out_rowmap[t]`` of ``out`` (``None`` = t).  multimodal_llama.py:295-312 without the scores."""
    qm, km, vm, om = _mat(q, "q"), _mat(k, "k", q.dtype), _mat(v, "v", q.dtype), _mat(out, "out", q.dtype)
    T = batch * seq_len
    if any(t.shape[0] != T for t in (qm, km, vm, om)) or qm.shape[1] != n_heads * 128 or not (qm.stride(0) == km.stride(0) == vm.stride(0)):
        raise ValueError("attention: q / k / v / out must be [batch * seq_len, n_heads * 128] with one common row stride")
    if out_rowmap is not None and (out_rowmap.dtype != torch.int32 or out_rowmap.numel() != T or not out_rowmap.is_cuda
                                   or not out_rowmap.is_contiguous()):
        raise ValueError("attention: out_rowmap must be a contiguous CUDA int32 [batch * seq_len]")
    _cabi.check(_cabi.lib().mc_attention_causal_tuned(qm.data_ptr(), km.data_ptr(), vm.data_ptr(), om.data_ptr(), qm.stride(0),
                                                      om.stride(0), None if out_rowmap is None else out_rowmap.data_ptr(), batch,
                                                      seq_len, n_heads, 128, float(softmax_scale), _cabi.dtype_code(q.dtype),
                                                      int(tuning), _cabi.current_stream_ptr()), "mc_attention_causal")
    _cabi.count_launch()
    return out


def gather_rows(src: torch.Tensor, index: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``out[i] = src[index[i]]`` over 2-D row-major views (bit-exact row copy on the GPU)."""
    s, o = _mat(src, "src"), _mat(out, "out", src.dtype)
    if index.dtype != torch.int32 or not index.is_cuda or not index.is_contiguous() or index.numel() != o.shape[0] or s.shape[1] != o.shape[1]:
        raise ValueError("gather_rows: index must be a contiguous CUDA int32 vector with one entry per output row")
    es = s.element_size()
    _cabi.check(_cabi.lib().mc_gather_rows(s.data_ptr(), s.stride(0) * es, o.data_ptr(), o.stride(0) * es, index.data_ptr(),
                                           o.shape[0], s.shape[1] * es, _cabi.current_stream_ptr()), "mc_gather_rows")
    _cabi.count_launch()
    return out


def silu_mul(gate: torch.Tensor, up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``silu(gate) * up`` over 2-D row-major views (multimodal_llama.py:381-388)."""
    g, u = _mat(gate, "gate"), _mat(up, "up", gate.dtype)
    if out is None:
        out = torch.empty(g.shape, dtype=g.dtype, device=g.device)
    o = _mat(out, "out", g.dtype)
    _cabi.check(_cabi.lib().mc_silu_mul(g.data_ptr(), u.data_ptr(), o.data_ptr(), g.shape[0], g.shape[1], g.stride(0),
                                        u.stride(0), o.stride(0), _cabi.dtype_code(g.dtype), _cabi.current_stream_ptr()),
                "mc_silu_mul")
    _cabi.count_launch()
    return out


# ------------------------------------------------------------------------------------------------ adapter packing
def pad64(n: int) -> int:
    return (n + K_BLOCK - 1) // K_BLOCK * K_BLOCK


@dataclass
class PackedAdapters:
    """Adapters of ONE LocalLoraLinear packed for the routed kernels.

    Group g = 0 is the text ("default") group, group 1 + i is ``modal_names[1 + i]`` (``infer_modals`` order, which
    is also the order of ``modal_id`` the splice emits).  The default group concatenates its sub-adapters
    ``default-{modal}`` (merged checkpoint, reset coefficients folded into the scaling, multimodal_llama.py:93-106,
    :130-149) or holds the single ``default`` adapter.  Every group's rank is padded to a multiple of 64 with zeros.

    ``A_all`` [R, in]   rows of group g = its A matrices stacked          (LoRA down: T = x·A_allᵀ, masked + scaled)
    ``B_all`` [out, R]  cols of group g = its B matrices side by side     (LoRA up:   y = x·Wᵀ + T·B_allᵀ)
    ``col_scale`` fp32 [R]  adapter scaling per rank column (0 on padding); applied to T in fp32 before rounding."""
    A_all: torch.Tensor
    B_all: torch.Tensor
    col_scale: torch.Tensor
    group_cols: List[int]


def pack_adapters(lora_A: Dict[str, torch.Tensor], lora_B: Dict[str, torch.Tensor], scaling: Dict[str, float],
                  modal_names: Sequence[str], default_adapter_names: Optional[Sequence[str]], in_features: int,
                  out_features: int, dtype, device) -> PackedAdapters:
    """Pure data movement (cat / zero-pad) at load time; ``modal_names[0]`` must be ``'default'``."""
    groups: List[List[str]] = []
    for name in modal_names:
        if name == "default" and default_adapter_names is not None:
            groups.append([n for n in default_adapter_names if n in lora_A])
        else:
            groups.append([name] if name in lora_A else [])
    cols = [0]
    for members in groups:
        r = sum(lora_A[n].shape[0] for n in members)
        cols.append(cols[-1] + pad64(r))
    R = max(cols[-1], K_BLOCK)
    if cols[-1] == 0:
        cols[-1] = R
    A_all = torch.zeros((R, in_features), dtype=dtype, device=device)
    B_all = torch.zeros((out_features, R), dtype=dtype, device=device)
    scale = torch.zeros(R, dtype=torch.float32, device=device)
    for g, members in enumerate(groups):
        c = cols[g]
        for n in members:
            r = lora_A[n].shape[0]
            A_all[c:c + r].copy_(lora_A[n])
            B_all[:, c:c + r].copy_(lora_B[n])
            scale[c:c + r] = float(scaling[n])
            c += r
    return PackedAdapters(A_all, B_all, scale, cols)
