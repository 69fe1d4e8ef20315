#!/usr/bin/env python
"""ncu target: a few launches of the tcgen05 causal attention kernel at one prefill shape (default C4: B=8, S=3046)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402

B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 3046)
nH, D = 32, 128
q, k, v = (torch.randn((B * S, nH * D), device="cuda", dtype=torch.bfloat16) for _ in range(3))
out = torch.empty_like(q)
for _ in range(3):
    LN.attention_causal(q, k, v, out, B, S, nH, 1.0 / math.sqrt(D))
torch.cuda.synchronize()
print("done")
