#!/usr/bin/env python
"""ncu target: the routed linears of ONE decoder layer at BASELINE config C3 shapes (31,360 rows), launched a few times.
Development / profiling aid, not a bench value.  Launch order per repetition: down_qkv, up_qkv, down_o, up_o, down_gu,
up_gu, down_d, up_d (8 linear_kernel launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import model as MD  # noqa: E402
from modelcompose_b200 import synthetic as syn  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda")
    cfg, base, sd = syn.make_composed_on_device(["audio", "vision", "video"], dev, torch.bfloat16, seed=1, layers=1)
    model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfg), base, sd, device=dev, dtype=torch.bfloat16)
    B, S = 32, 980
    x = torch.randn(B, S, 4096, device=dev, dtype=torch.bfloat16) * 0.5
    seg = torch.zeros(S, dtype=torch.uint8)
    seg[40:40 + 586] = 2   # vision block (+5+5)
    seg[630:630 + 266] = 1  # audio block
    mid = seg[None].expand(B, S).contiguous().to(dev)
    for _ in range(reps):
        model.prefill(x, mid, None)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
