#!/usr/bin/env python
"""Development aid: which stock-library causal attention is fastest on this B200 for the prefill shapes."""
import math
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for (B, S) in [(32, 980), (8, 3046), (16, 3569)]:
    nH, D = 32, 128
    q, k, v = (torch.randn(B, S, nH, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    flops = 4.0 * B * nH * S * S * D / 2
    res = {}
    try:
        from flash_attn import flash_attn_func
        ref = flash_attn_func(q, k, v, causal=True)
        res["flash_attn2"] = timeit(lambda: flash_attn_func(q, k, v, causal=True))
    except Exception as e:
        print("flash_attn failed", e)
        ref = None
    for name, be in (("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash", SDPBackend.FLASH_ATTENTION),
                     ("sdpa_efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel(be):
                o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
                if ref is not None:
                    err = (o.transpose(1, 2).float() - ref.float()).abs().max().item()
                else:
                    err = float("nan")
                res[name] = timeit(lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True))
                print(f"   {name}: max diff vs flash_attn2 {err:.4g}, output contiguous-as-[B,S,H,D]: {o.transpose(1, 2).is_contiguous()}")
        except Exception as e:
            print(f"   {name} unavailable: {str(e)[:120]}")
    print(f"B={B} S={S}: " + ", ".join(f"{k_} {v_:.3f} ms ({flops / v_ / 1e9:.0f} TF/s)" for k_, v_ in res.items()), flush=True)
