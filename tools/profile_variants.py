#!/usr/bin/env python
"""ncu target: a few launches of chosen linear-kernel variants on chosen shapes.
  python tools/profile_variants.py "M,N,K;M,N,K" "tuning[:group_m],..." [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402

shapes = [tuple(int(x) for x in s.split(",")) for s in sys.argv[1].split(";")]
variants = [tuple(int(x) for x in (v.split(":") + ["0"])[:2]) for v in sys.argv[2].split(",")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for M, N, K in shapes:
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for tuning, gm in variants:
        plan = LN.LinearPlan([LN.Problem(A, B, C)], tuning=tuning | (gm << 8))
        for _ in range(iters):
            plan.run()
        torch.cuda.synchronize()
        if os.environ.get("MC_TIME"):  # CUDA-event timing instead of an ncu target
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = int(os.environ["MC_TIME"])
            a.record()
            for _ in range(n):
                plan.run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n
            print(f"M={M} N={N} K={K} tuning={tuning} group_m={gm or 'default'}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
print("done")
