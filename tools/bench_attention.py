#!/usr/bin/env python
"""Time the tcgen05 causal attention kernel against the stock cuDNN call at the prefill shapes (CUDA events, 20 runs)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402

F = torch.nn.functional
for B, S in ((32, 980), (8, 3046), (16, 3569)):
    nH, D = 32, 128
    T, H = B * S, nH * D
    q, k, v = (torch.randn((T, H), device="cuda", dtype=torch.bfloat16) for _ in range(3))
    out = torch.empty_like(q)
    scale = 1.0 / math.sqrt(D)
    flops = 4.0 * B * nH * D * S * (S + 1) / 2  # causal: QK^T and PV over the lower triangle

    def native():
        LN.attention_causal(q, k, v, out, B, S, nH, scale)

    def cudnn():
        from torch.nn.attention import SDPBackend, sdpa_kernel
        with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
            return F.scaled_dot_product_attention(q.view(B, S, nH, D).transpose(1, 2), k.view(B, S, nH, D).transpose(1, 2),
                                                  v.view(B, S, nH, D).transpose(1, 2), is_causal=True, scale=scale)
    for name, fn in (("native", native), ("cudnn", cudnn)):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"B={B} S={S}: {name} {ms:.3f} ms ({flops / ms / 1e9:.0f} TF/s causal)", flush=True)
    ref = cudnn().transpose(1, 2).reshape(T, H)
    native()
    torch.cuda.synchronize()
    print(f"   max |native - cudnn| = {(out.float() - ref.float()).abs().max().item():.4f} (max |ref| {ref.float().abs().max().item():.3f})", flush=True)
