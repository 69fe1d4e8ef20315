#!/usr/bin/env python
"""ncu target: one 8192^3 product on the CTA-pair kernel (tuning 3) and on the single-CTA kernel (tuning 2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402

M = N = K = 8192
A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for tuning in (3, 2):
    plan = LN.LinearPlan([LN.Problem(A, B, C)], tuning=tuning)
    for _ in range(3):
        plan.run()
torch.cuda.synchronize()
print("done")
