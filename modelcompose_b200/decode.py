"""Decode step of the composed model: one new token per sequence against the key/value cache, as one CUDA graph.

Reference semantics (all with ``past_key_values`` present):
  modelcompose/model/multimodal_arch.py:290-293                      no splice; the attention mask is rebuilt over past + 1
  modelcompose/model/language_model/multimodal_llama.py:436-438      the modality masks are dropped: every row takes the default
                                                                     adapter (group 0 = ``default`` or the ``default-{modal}`` set)
  :262-312 attention with the cached keys, :380-390 MLP, :720 lm_head, :747-767 prepare_inputs_for_generation
and the greedy loop HF ``generate`` runs for modelcompose/eval/model_multimodal_qa_loader.py:93-102.

With at most 64 rows every linear is one pass over its weight matrix: the step is HBM-bound on weight + cache bytes, and runs
on the kernels of csrc/mc_decode.cu (``mc_skinny_linear``, ``mc_decode_rope_append``, ``mc_decode_attention``,
``mc_argmax_rows``) instead of the prefill's 128-row tcgen05 tiles.  The whole step — embedding gather, 32 layers, final norm,
lm_head, greedy argmax feeding the next step's ids, position counter — is captured once per (batch, cache) into a CUDA graph
and replayed: no host synchronisation and no Python launch overhead between tokens.  There is no torch fallback.
"""
from __future__ import annotations

import ctypes as C
import gc
import math
import os
from typing import List, Optional

import torch

from . import _cabi
from . import linear as LN

SK_NONE, SK_RESIDUAL, SK_COLSCALE, SK_SILU_MUL = 0, 1, 2, 3
# tuning bits OR-ed into the low-rank "down" launches of the branch form (3 - 9 MB of weights each): 16 = the register kernel
DOWN_TUNING = int(os.environ.get("MC_DECODE_DOWN_TUNING", "0"))
MAX_M = 64
# programmatic dependent launch along the decode chain (mc_set_launch_mode): every kernel of the step may start while its
# predecessor still runs and the skinny linears prefetch their first ring of weights before they wait for it; 0 = plain launches
PDL = os.environ.get("MC_DECODE_PDL", "1") != "0"
# RoPE of the new q / k and the cache append inside the attention launch (mc_decode_attention_fused); 0 = two launches (A/B switch)
FUSED_ROPE = os.environ.get("MC_DECODE_FUSED_ROPE", "1") != "0"
# 1 = the RMSNorms of the step inside the skinny launch that consumes them (mc_skinny_plan_set_norm: bit-identical, 65 launches fewer).
# Default 0: measured 2-3 % SLOWER (7.50 vs 7.29 ms materialised, profiles/r02_decode.txt step 11) — with programmatic dependent launch the
# separate rmsnorm kernel already runs under the next launch's weight prefetch, while the in-kernel norm puts a cross-CTA wait in front
# of the activation loads.
FUSED_NORM = os.environ.get("MC_DECODE_FUSED_NORM", "0") != "0"


class SkinnyDesc(C.Structure):
    """mirror of mc_skinny_desc_t"""
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K0", C.c_int32), ("K1", C.c_int32),
                ("A0", C.c_void_p), ("lda0", C.c_int64), ("B0", C.c_void_p), ("ldb0", C.c_int64),
                ("A1", C.c_void_p), ("lda1", C.c_int64), ("B1", C.c_void_p), ("ldb1", C.c_int64),
                ("B0u", C.c_void_p), ("A1u", C.c_void_p), ("B1u", C.c_void_p),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("residual", C.c_void_p), ("ldr", C.c_int64),
                ("col_scale", C.c_void_p), ("epilogue", C.c_int32)]


def _m(t: torch.Tensor, what: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{what} must be a 2-D CUDA tensor with unit stride along the last dimension")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{what}: expected {dtype}, got {t.dtype}")
    return t


_WORKSPACES = {}


def skinny_workspace(device) -> torch.Tensor:
    """Zeroed scratch of the stream-K kernel (one per device and stream user; the kernel leaves it zeroed)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _WORKSPACES:
        n = int(_cabi.lib().mc_skinny_workspace_bytes())
        _WORKSPACES[key] = torch.zeros(n, dtype=torch.uint8, device=f"cuda:{key}")
    return _WORKSPACES[key]


class SkinnyLaunch:
    """1..4 problems ``C = epilogue(A0·B0ᵀ + A1·B1ᵀ)`` with M <= 64 rows: a ``mc_skinny_plan`` (TMA descriptors + stream-K
    schedule), one launch per ``run``."""

    def __init__(self, problems: List[dict], tuning: int = 0, workspace: Optional[torch.Tensor] = None):
        if not 1 <= len(problems) <= 4:
            raise ValueError("1..4 problems per launch")
        self.dtype = problems[0]["A0"].dtype
        self.descs = (SkinnyDesc * len(problems))()
        self.keep = problems
        self.tuning = int(tuning)
        self.bytes = 0
        for d, p in zip(self.descs, problems):
            A0, B0, Cm = _m(p["A0"], "A0", self.dtype), _m(p["B0"], "B0", self.dtype), _m(p["C"], "C", self.dtype)
            M, K0 = A0.shape
            N = B0.shape[0]
            if B0.shape[1] != K0 or tuple(Cm.shape) != (M, N) or M > MAX_M:
                raise ValueError(f"skinny linear: shapes A0 {tuple(A0.shape)} B0 {tuple(B0.shape)} C {tuple(Cm.shape)} (M <= {MAX_M})")
            d.M, d.N, d.K0, d.K1 = M, N, K0, 0
            d.A0, d.lda0, d.B0, d.ldb0, d.C, d.ldc = A0.data_ptr(), A0.stride(0), B0.data_ptr(), B0.stride(0), Cm.data_ptr(), Cm.stride(0)
            d.epilogue = int(p.get("epilogue", SK_NONE))
            w_elems = N * K0
            if p.get("A1") is not None:
                A1, B1 = _m(p["A1"], "A1", self.dtype), _m(p["B1"], "B1", self.dtype)
                if A1.shape[0] != M or B1.shape[0] != N or A1.shape[1] != B1.shape[1]:
                    raise ValueError("skinny linear: A1 / B1 do not match M, N")
                d.K1, d.A1, d.lda1, d.B1, d.ldb1 = A1.shape[1], A1.data_ptr(), A1.stride(0), B1.data_ptr(), B1.stride(0)
                w_elems += N * A1.shape[1]
            if d.epilogue == SK_SILU_MUL:
                B0u = _m(p["B0u"], "B0u", self.dtype)
                if tuple(B0u.shape) != tuple(B0.shape) or B0u.stride(0) != B0.stride(0):
                    raise ValueError("skinny linear: B0u must match B0")
                d.B0u = B0u.data_ptr()
                w_elems += N * K0
                if d.K1:
                    A1u, B1u = _m(p["A1u"], "A1u", self.dtype), _m(p["B1u"], "B1u", self.dtype)
                    if tuple(A1u.shape) != tuple(p["A1"].shape) or A1u.stride(0) != p["A1"].stride(0) \
                            or tuple(B1u.shape) != tuple(p["B1"].shape) or B1u.stride(0) != p["B1"].stride(0):
                        raise ValueError("skinny linear: A1u / B1u must match A1 / B1 in shape and stride")
                    d.A1u, d.B1u = A1u.data_ptr(), B1u.data_ptr()
                    w_elems += N * d.K1
            if p.get("residual") is not None:
                R = _m(p["residual"], "residual", self.dtype)
                d.residual, d.ldr = R.data_ptr(), R.stride(0)
            if p.get("col_scale") is not None:
                cs = p["col_scale"]
                if cs.dtype != torch.float32 or cs.numel() != N or not cs.is_cuda or not cs.is_contiguous():
                    raise ValueError("skinny linear: col_scale must be contiguous CUDA fp32 [N]")
                d.col_scale = cs.data_ptr()
            self.bytes += 2 * w_elems  # weight bytes streamed (the roofline's numerator)
        self.ws = workspace if workspace is not None else skinny_workspace(problems[0]["A0"].device)
        self._h = C.c_void_p()
        _cabi.check(_cabi.lib().mc_skinny_plan_create(C.byref(self._h), self.descs, len(self.descs), _cabi.dtype_code(self.dtype),
                                                      self.tuning), "mc_skinny_plan_create")
        assert int(_cabi.lib().mc_skinny_plan_bytes(self._h)) == self.bytes

    def set_norm(self, src: torch.Tensor, weight: torch.Tensor, dst: torch.Tensor, eps: float) -> None:
        """The launch first computes ``dst = rmsnorm(src) * weight`` (the activations its problems read) under its weight ramp."""
        s_, d_ = _m(src, "norm src", self.dtype), _m(dst, "norm dst", self.dtype)
        if tuple(s_.shape) != tuple(d_.shape) or weight.dtype != self.dtype or weight.numel() != s_.shape[1] or not weight.is_cuda:
            raise ValueError("skinny linear: RMSNorm src / dst / weight do not match")
        self.keep_norm = (src, weight, dst)
        _cabi.check(_cabi.lib().mc_skinny_plan_set_norm(self._h, s_.data_ptr(), s_.stride(0), weight.data_ptr(), d_.data_ptr(), d_.stride(0),
                                                        s_.shape[0], s_.shape[1], float(eps)), "mc_skinny_plan_set_norm")

    def run(self) -> None:
        _cabi.check(_cabi.lib().mc_skinny_plan_run(self._h, self.ws.data_ptr(), self.ws.numel(), _cabi.current_stream_ptr()),
                    "mc_skinny_plan_run")
        _cabi.count_launch()

    def close(self) -> None:
        if getattr(self, "_h", None):
            _cabi.lib().mc_skinny_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def argmax_rows(logits: torch.Tensor, out_i32: Optional[torch.Tensor] = None, out_i64: Optional[torch.Tensor] = None,
                counter: Optional[torch.Tensor] = None) -> None:
    lg = _m(logits, "logits")
    _cabi.check(_cabi.lib().mc_argmax_rows(lg.data_ptr(), lg.stride(0), lg.shape[0], lg.shape[1],
                                           None if out_i32 is None else out_i32.data_ptr(),
                                           None if out_i64 is None else out_i64.data_ptr(),
                                           None if counter is None else counter.data_ptr(),
                                           _cabi.dtype_code(lg.dtype), _cabi.current_stream_ptr()), "mc_argmax_rows")
    _cabi.count_launch()


class DecodeWorkspace:
    """Buffers, launches and the captured graph of the decode step for one (model, batch, key/value cache)."""

    def __init__(self, model, cache, key_mask: Optional[torch.Tensor] = None, tuning: int = 0, use_graph: bool = True):
        cfg, dev, dt = model.config, model.device, model.dtype
        B = cache.k[0].shape[0]  # caches: [B, heads, capacity, head_dim]
        H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
        nH = cfg.num_attention_heads
        D = H // nH
        if B > MAX_M or D != 128:
            raise ValueError(f"decode kernels take batch <= {MAX_M} and head_dim 128")
        self.model, self.cache, self.B, self.nH, self.D = model, cache, B, nH, D
        self.capacity = cache.capacity
        self.cache_ptrs = tuple(t.data_ptr() for t in cache.k)
        self.key_mask = key_mask  # uint8 [B, capacity] or None
        self.use_graph = use_graph
        self.pdl = PDL
        self.fused_rope = FUSED_ROPE
        self.fused_norm = FUSED_NORM
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.warm = 0

        def buf(*shape, dtype=dt):
            return torch.zeros(shape, dtype=dtype, device=dev)
        self.ids = buf(B, dtype=torch.int32)        # token ids of the step (the in-graph argmax writes the next step's here)
        self.next64 = buf(B, dtype=torch.int64)
        self.pos = buf(1, dtype=torch.int32)        # tokens cached so far = position of the new token
        self.x, self.xn = buf(B, H), buf(B, H)
        self.q, self.k, self.v, self.attn = buf(B, H), buf(B, H), buf(B, H), buf(B, H)
        self.act = buf(B, I)
        self.logits = buf(B, V)
        # split the keys of a (sequence, head) over CTAs so that ONE wave fills the GPU: 7 CTAs of the attention kernel fit an SM
        # (72 registers x 128 threads), and a count just under 7 x SMs measured best (B = 32: 1 split, 16: 2, 8: 3-4, 1: 32)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.n_splits = max(1, min(32, (7 * sms) // (B * nH), max(1, self.capacity // 64)))
        self.att_scratch = buf(B * nH * self.n_splits * (D + 2), dtype=torch.float32)
        self.att_counters = buf(B * nH, dtype=torch.int32)
        ws = skinny_workspace(dev)
        dense = model.materialize or getattr(model, "decode_dense", False)
        R0 = 0 if dense else int(model.layers[0].ad["q_proj"].group_cols[1])  # rank columns of the default group
        self.t = [buf(B, max(R0, 8)) for _ in range(3)] if R0 else []
        self.launches: List[List] = []
        self.weight_bytes = 0

        def W(layer, n):
            return layer.Weff[n][0] if model.materialize else (layer.Wdec[n] if dense else layer.W[n])

        def down(layer, names, src):
            probs = [dict(A0=src, B0=layer.ad[n].A_all[:R0], C=self.t[i], col_scale=layer.ad[n].col_scale[:R0].contiguous(),
                          epilogue=SK_COLSCALE) for i, n in enumerate(names)]
            return SkinnyLaunch(probs, tuning | DOWN_TUNING, ws)

        def up(layer, names, src, outs, residual=None, tune=tuning):
            probs = []
            for i, (n, o) in enumerate(zip(names, outs)):
                p = dict(A0=src, B0=W(layer, n), C=o, epilogue=SK_RESIDUAL if residual is not None else SK_NONE, residual=residual)
                if R0:
                    p.update(A1=self.t[i], B1=layer.ad[n].B_all[:, :R0])
                probs.append(p)
            return SkinnyLaunch(probs, tune, ws)

        for layer in model.layers:
            L = {}
            if R0:
                L["down_qkv"] = down(layer, ("q_proj", "k_proj", "v_proj"), self.xn)
                L["down_o"] = down(layer, ("o_proj",), self.attn)
                L["down_gu"] = down(layer, ("gate_proj", "up_proj"), self.xn)
                L["down_d"] = down(layer, ("down_proj",), self.act)
            L["qkv"] = up(layer, ("q_proj", "k_proj", "v_proj"), self.xn, (self.q, self.k, self.v))
            # o_proj (33.5 MB): the register kernel beats the stream-K kernel on launches this small (16.5 vs 22 us at M = 32: every CTA
            # of the stream-K schedule would share both its row blocks with a neighbour), profiles/r02_decode.txt
            L["o"] = up(layer, ("o_proj",), self.attn, (self.x,), residual=self.x, tune=tuning | 16)
            gu = dict(A0=self.xn, B0=W(layer, "gate_proj"), B0u=W(layer, "up_proj"), C=self.act, epilogue=SK_SILU_MUL)
            if R0:
                gu.update(A1=self.t[0], B1=layer.ad["gate_proj"].B_all[:, :R0], A1u=self.t[1], B1u=layer.ad["up_proj"].B_all[:, :R0])
            L["gu"] = SkinnyLaunch([gu], tuning, ws)
            L["d"] = up(layer, ("down_proj",), self.act, (self.x,), residual=self.x)
            if self.fused_norm:  # the first launch that reads xn computes it: input_layernorm -> q/k/v side, post_attention_layernorm -> MLP side
                eps = float(cfg.rms_norm_eps)
                L["down_qkv" if R0 else "qkv"].set_norm(self.x, layer.ln1, self.xn, eps)
                L["down_gu" if R0 else "gu"].set_norm(self.x, layer.ln2, self.xn, eps)
            self.launches.append(L)
            self.weight_bytes += sum(v.bytes for v in L.values())
        self.lm_head = SkinnyLaunch([dict(A0=self.xn, B0=model.lm_head, C=self.logits)], tuning, ws)
        if self.fused_norm:
            self.lm_head.set_norm(self.x, model.norm, self.xn, float(cfg.rms_norm_eps))
        self.weight_bytes += self.lm_head.bytes

    # ---------------------------------------------------------------------------------------------------------------
    def matches(self, cache) -> bool:
        return cache is self.cache and cache.capacity == self.capacity and tuple(t.data_ptr() for t in cache.k) == self.cache_ptrs

    def cache_bytes(self, length: int) -> int:
        """Key/value bytes one step reads at ``length`` cached tokens (all layers)."""
        return 2 * len(self.launches) * self.B * length * self.nH * self.D * self.x.element_size()

    def _attention(self, li: int) -> None:
        m, lib, dtc, st = self.model, _cabi.lib(), _cabi.dtype_code(self.model.dtype), _cabi.current_stream_ptr()
        cos, sin = m._rope
        kc, vc = self.cache.k[li], self.cache.v[li]
        km = self.key_mask
        if self.fused_rope:
            _cabi.check(lib.mc_decode_attention_fused(self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), self.q.stride(0), kc.data_ptr(),
                                                      vc.data_ptr(), self.capacity, self.pos.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                                      None if km is None else km.data_ptr(), 0 if km is None else km.stride(0),
                                                      self.attn.data_ptr(), self.attn.stride(0), self.B, self.nH, self.D,
                                                      1.0 / math.sqrt(self.D), self.n_splits, self.att_scratch.data_ptr(),
                                                      self.att_counters.data_ptr(), dtc, st), "mc_decode_attention_fused")
            _cabi.count_launch()
            return
        _cabi.check(lib.mc_decode_rope_append(self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), self.q.stride(0), kc.data_ptr(),
                                              vc.data_ptr(), self.capacity, self.pos.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                              self.B, self.nH, self.D, dtc, st), "mc_decode_rope_append")
        _cabi.check(lib.mc_decode_attention(self.q.data_ptr(), kc.data_ptr(), vc.data_ptr(), self.capacity, self.pos.data_ptr(),
                                            None if km is None else km.data_ptr(), 0 if km is None else km.stride(0),
                                            self.attn.data_ptr(), self.q.stride(0), self.attn.stride(0), self.B, self.nH, self.D,
                                            1.0 / math.sqrt(self.D), self.n_splits, self.att_scratch.data_ptr(),
                                            self.att_counters.data_ptr(), dtc, st), "mc_decode_attention")
        _cabi.count_launch(2)

    def _enqueue(self) -> None:
        """All launches of one step on the current stream (this is what the graph records)."""
        prev = _cabi.lib().mc_set_launch_mode(1 if self.pdl else 0)
        try:
            self._enqueue_chain()
        finally:
            _cabi.lib().mc_set_launch_mode(prev)

    def _enqueue_chain(self) -> None:
        m = self.model
        LN.gather_rows(m.embed_tokens, self.ids, self.x)
        for li, (layer, L) in enumerate(zip(m.layers, self.launches)):
            if not self.fused_norm:
                m._rmsnorm(self.x, layer.ln1, self.xn)
            if "down_qkv" in L:
                L["down_qkv"].run()
            L["qkv"].run()
            self._attention(li)
            if "down_o" in L:
                L["down_o"].run()
            L["o"].run()
            if not self.fused_norm:
                m._rmsnorm(self.x, layer.ln2, self.xn)
            if "down_gu" in L:
                L["down_gu"].run()
            L["gu"].run()
            if "down_d" in L:
                L["down_d"].run()
            L["d"].run()
        if not self.fused_norm:
            m._rmsnorm(self.x, m.norm, self.xn)
        self.lm_head.run()
        # greedy sampler + hand-over: next ids into the gather index of the next step, position counter + 1
        argmax_rows(self.logits, self.ids, self.next64, self.pos)

    def launches_per_step(self) -> int:
        per_layer = len(self.launches[0]) + (1 if self.fused_rope else 2) + (0 if self.fused_norm else 2) if self.launches else 0
        return 1 + per_layer * len(self.launches) + (2 if self.fused_norm else 3)

    def run(self) -> None:
        """One decode step from the state in ``ids`` / ``pos``; leaves logits, next ids (``ids`` / ``next64``) and ``pos`` + 1."""
        if not self.use_graph:
            self._enqueue()
            return
        if self.graph is None:
            if self.warm < 1:
                # first step eagerly: lazy module loading and func attributes must not happen under capture
                self.warm += 1
                self._enqueue()
                return
            g = torch.cuda.CUDAGraph()
            n0 = _cabi.LAUNCHES
            # capture WITHOUT executing: the recorded step is replayed right away.  A cudaFree invalidates a capture in progress, and
            # the finalizers of plan objects (this package's and anyone's) issue one whenever the garbage collector decides to run
            # them: collect what is pending first, keep the collector off for the few milliseconds of the capture, and let other
            # threads' CUDA calls be (thread_local error mode).
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._enqueue()
            finally:
                if gc_was_on:
                    gc.enable()
            _cabi.LAUNCHES = n0
            self.graph = g
        self.graph.replay()
        _cabi.count_launch(self.launches_per_step())

    def step(self, ids: torch.Tensor, length: int) -> torch.Tensor:
        """Decode step for explicit token ids [B] (int) at ``length`` cached tokens; returns the logits buffer [B, V]."""
        self.ids.copy_(ids.reshape(-1))
        self.pos.fill_(int(length))
        self.run()
        return self.logits
