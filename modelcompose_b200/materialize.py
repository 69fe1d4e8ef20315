"""Materialised effective weights of a composed model: ``W_eff,g = W + Σ_a s_a · B_a · A_a`` per routing group.

The reference evaluates the blend per forward, as extra low-rank branches on every token (``LocalLoraLinear.forward``,
modelcompose/model/language_model/multimodal_llama.py:130-149, coefficients :93-106); its own tooling writes the same thing as
dense weights (``scripts/model_composition/delta_weights_compare.py:24-31,61`` — ``W + (B @ A) * scale`` — and
``scripts/convert_to_multimodal.py:111-113``).  Here the dense form is built ON THE DEVICE with the library's kernels:

  * every modality group g >= 1:   W_eff,g = W + s_g B_g A_g                                  (one rank-r GEMM, residual epilogue)
  * the text ("default") group:    D_m = W + s B_{default-m} A_{default-m}  for every merged modality m — the dense unimodal
    checkpoints the composition started from — and then the ONLINE-MERGE-RESET blend of those checkpoints,
        W_eff,0 = (1 − Σ_m w_m) · W + Σ_m w_m · D_m,
    by the N-source merge kernel (``mc_merge_plan_*``, MC_MERGE_WEIGHTED): the 3 x 7B merge of BASELINE config 2 with the base
    as a fourth source.  w_m are the reset coefficients of the merge CLI's ``--strategy online-merge-reset-default-<m>=w_m``.

No arithmetic runs in torch: the adapter scalings are applied by the merge kernel (one source, weight s: an exact fp32 product
rounded once), the products by the tcgen05 linear kernel, the blend by the merge kernel.  torch only transposes / allocates.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import linear as LN
from . import merge as MG


def _scaled_copies(tensors: Sequence[torch.Tensor], scale: float) -> List[torch.Tensor]:
    """``rn(scale * t)`` for every tensor, one multi-tensor launch of the merge kernel."""
    outs = [torch.empty_like(t) for t in tensors]
    if tensors:
        plan = MG.MergePlan([list(tensors)], outs)
        plan.run([scale])
        plan.close()
    return outs


def dense_plus_lora(W: torch.Tensor, A_scaled_T: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``W + B · A_scaled`` with A_scaled given transposed ([in, r], K-major for the tensor core); fp32 accumulation, one rounding."""
    out = torch.empty_like(W) if out is None else out
    LN.LinearPlan([LN.Problem(B, A_scaled_T, out, residual=W, epilogue=LN.EPI_RESIDUAL)]).run()
    return out


def effective_weights(W: torch.Tensor, lora_A: Dict[str, torch.Tensor], lora_B: Dict[str, torch.Tensor],
                      scaling: Dict[str, float], modal_names: Sequence[str], default_adapter_names: Optional[Sequence[str]],
                      base_scale: float, reset: Optional[Dict[str, float]] = None) -> List[torch.Tensor]:
    """One dense ``[out, in]`` weight per routing group of ONE linear (group 0 = text, then ``modal_names[1:]``).

    ``scaling[a]`` is the effective scaling of adapter a (reset coefficient folded in, multimodal_llama.py:98-106);
    ``base_scale`` = lora_alpha / r and ``reset`` the coefficient dict, so that the text group can be built as the blend of
    the unimodal dense checkpoints (see the module docstring).  Groups without adapter weights reuse ``W`` itself."""
    dtype = W.dtype
    out: List[torch.Tensor] = []
    for gi, name in enumerate(modal_names):
        if gi == 0 and default_adapter_names is not None:
            members = [n for n in default_adapter_names if n in lora_A]
            if not members:
                out.append(W)
                continue
            # dense unimodal checkpoints D_m (adapter scaling lora_alpha / r, before the reset coefficient)
            A_s = _scaled_copies([lora_A[n] for n in members], base_scale)
            dense = [dense_plus_lora(W, a.t().contiguous(), lora_B[n]) for a, n in zip(A_s, members)]
            w = [float((reset or {}).get(n, scaling[n] / base_scale)) for n in members]
            Weff = torch.empty_like(W)
            plan = MG.MergePlan([[W]] + [[d] for d in dense], [Weff])
            plan.run([1.0 - sum(w)] + w)   # (1 − Σw)·W + Σ w_m·D_m — the online-merge-reset blend of N checkpoints
            plan.close()
            del dense
            out.append(Weff)
        elif name in lora_A:
            (a_s,) = _scaled_copies([lora_A[name]], float(scaling[name]))
            out.append(dense_plus_lora(W, a_s.t().contiguous(), lora_B[name]))
        else:
            out.append(W)
    assert all(t.dtype == dtype for t in out)
    return out
