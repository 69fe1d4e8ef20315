#!/usr/bin/env python
"""Time the TIES merge (all passes of mc_ties_plan_run) on one GPU: CUDA events around `iters` runs, algorithmic GB/s.

  python tools/bench_ties.py [--elements 160e6] [--src 3] [--dtype bf16] [--func mean] [--K 20] [--kind gauss|zeros|neg]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import merge as M  # noqa: E402

DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elements", type=float, default=160e6)
    ap.add_argument("--src", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=sorted(DT))
    ap.add_argument("--func", default="mean")
    ap.add_argument("--K", type=float, default=20)
    ap.add_argument("--kind", default="gauss")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    n = int(args.elements)
    dt = DT[args.dtype]
    g = torch.Generator(device="cuda").manual_seed(1)
    sizes = [n // 2, n // 4, n - n // 2 - n // 4]
    srcs = []
    for _ in range(args.src):
        lst = []
        for m in sizes:
            if args.kind == "zeros":
                t = torch.zeros(m, device="cuda")
            else:
                t = torch.randn(m, generator=g, device="cuda") * 0.02 - (0.03 if args.kind == "neg" else 0.0)
            lst.append(t.to(dt))
        srcs.append(lst)
    odt = torch.float32 if args.func == "mean" else dt
    outs = [torch.empty(m, dtype=odt, device="cuda") for m in sizes]
    plan = M.TiesPlan(srcs, outs)
    for _ in range(3):
        plan.run(args.K, args.func)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        plan.run(args.K, args.func)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.iters
    st = plan.stats()
    print(json.dumps({"elements": n, "n_src": args.src, "dtype": args.dtype, "func": args.func, "K": args.K, "kind": args.kind,
                      "ms": round(ms, 4), "algorithmic_bytes": plan.algorithmic_bytes,
                      "GBps": round(plan.algorithmic_bytes / ms / 1e6, 1), "stats": st}))


if __name__ == "__main__":
    main()
