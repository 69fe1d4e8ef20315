#!/usr/bin/env python
"""Tuning sweep of the merge kernel variants on one GPU (development aid; not a bench value).

Prints GB/s (algorithmic bytes / CUDA-event time, best and median of --iters) per tuning code, plus a plain
device copy measured the same way as the roofline denominator in MEASURED_PEAKS.json."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import merge as M  # noqa: E402
from modelcompose_b200 import synthetic as syn  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--nsrc", type=int, default=3)
    ap.add_argument("--tunings", default="0,1,2,3,4,5")
    ap.add_argument("--ctas", default="0")
    args = ap.parse_args()
    shapes = [s for n, s in syn.dense_7b_tensor_shapes() if n.startswith("model.layers.") and int(n.split(".")[2]) < args.layers]
    srcs = [[torch.randn(s, device="cuda", dtype=torch.bfloat16) * 0.02 for s in shapes] for _ in range(args.nsrc)]
    outs = [torch.empty(s, device="cuda", dtype=torch.bfloat16) for s in shapes]
    w = [1.0 / args.nsrc] * args.nsrc
    n = 1 << 30
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
    b = torch.empty_like(a)
    best, med = timeit(lambda: b.copy_(a), args.iters)
    print(f"torch copy 1Gi bf16: best {2 * n * 2 / best / 1e6:.1f} GB/s  median {2 * n * 2 / med / 1e6:.1f} GB/s", flush=True)
    del a, b
    for ctas in [int(x) for x in args.ctas.split(",")]:
        for base in [int(x) for x in args.tunings.split(",")]:
            for nonpers in (0, 1):
                tuning = (1 << 24) | base | (ctas << 8) | (nonpers << 16)
                plan = M.MergePlan(srcs, outs, tuning=tuning)
                best, med = timeit(lambda: plan.run(w), args.iters)
                gb = plan.algorithmic_bytes / 1e9
                print(f"tuning {base} ctas/sm {ctas or 'max'} {'grid=chunks' if nonpers else 'persistent'}: "
                      f"best {gb / best * 1e3:.1f} GB/s  median {gb / med * 1e3:.1f} GB/s  ({med:.3f} ms, {gb:.2f} GB)", flush=True)
                plan.close()


if __name__ == "__main__":
    main()
