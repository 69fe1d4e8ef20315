#!/usr/bin/env python
"""Development aid: time the splice (plan + gather) at BASELINE config shapes on one GPU.  Not a bench value."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import splice as SP  # noqa: E402

CONFIGS = {  # name: (batch per GPU, [(modal, rows)], text tokens)
    "c3": (32, [("vision", 576), ("audio", 256)], 128),
    "c4": (8, [("video", 2056), ("vision", 576), ("audio", 256)], 128),
    "c5": (16, [("vision", 576), ("audio", 256), ("video", 2056), ("point", 513)], 128),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    B, modal, text = CONFIGS[args.config]
    H, V = 4096, 32000
    g = torch.Generator(device="cuda").manual_seed(0)
    S = text + len(modal)
    ids = torch.randint(3, V, (B, S), generator=g, device="cuda")
    for i, (m, _) in enumerate(modal):
        ids[:, 36 + 3 * i] = SP.MODAL_TOKEN_INDEXES[m]
    embed = torch.randn(V, H, generator=g, device="cuda", dtype=torch.bfloat16)
    order = [m for m in ("audio", "vision", "video", "point") if m in dict(modal)]
    feats = {m: torch.randn(B, dict(modal)[m], H, generator=g, device="cuda", dtype=torch.bfloat16) for m in order}
    pre = {m: torch.randn(1, 5, H, generator=g, device="cuda", dtype=torch.bfloat16) for m in order}
    suf = {m: torch.randn(1, 5, H, generator=g, device="cuda", dtype=torch.bfloat16) for m in order}
    attn = torch.ones(B, S, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(args.iters + 3):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = SP.splice(ids, attn, None, embed, feats, pre, suf)
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{args.config}: rows={r.inputs_embeds.shape[0] * r.inputs_embeds.shape[1]} bytes={r.algorithmic_bytes / 1e6:.1f} MB "
          f"wall(plan+gather+sync) median {med * 1e6:.1f} us best {ts[0] * 1e6:.1f} us -> {r.algorithmic_bytes / med / 1e9:.1f} GB/s e2e")


if __name__ == "__main__":
    main()
