#!/usr/bin/env python
"""Standalone timing (CUDA events, 20 runs each) and ncu target of the linear kernel variants on prefill-shaped products:
single-CTA 128x256 tiles (tuning 2), CTA pair 512x256 (3), CTA pair 256x256 with overlapped epilogue (4); optional
rasterisation-group override (M tiles per N sweep) in tuning bits 8-15."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402

SHAPES = [(8192, 8192, 8192), (31360, 4096, 4096), (31360, 11008, 4096), (31360, 4096, 11008)]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
VARIANTS = [(2, 0), (2, 4), (2, 16), (3, 0), (3, 1), (3, 2), (3, 4), (4, 0), (4, 2), (4, 4), (4, 16)]
for M, N, K in SHAPES:
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ref = None
    for tuning, gm in VARIANTS:
        plan = LN.LinearPlan([LN.Problem(A, B, C)], tuning=tuning | (gm << 8))
        for _ in range(3):
            plan.run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            plan.run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        same = True if ref is None else bool(torch.equal(ref, C))
        if ref is None:
            ref = C.clone()
        print(f"M={M} N={N} K={K} tuning={tuning} group_m={gm or 'default'}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s  identical={same}", flush=True)
print("done")
