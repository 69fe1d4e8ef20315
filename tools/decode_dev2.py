"""Stream-K skinny kernel: where does the time go?  Sweeps the profiling knobs (copy-only consumers, ring depth, chunk size)
at vicuna-7B shapes.  python tools/decode_dev2.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import decode as DC  # noqa: E402

H, I, V = 4096, 11008, 32000
dt = torch.bfloat16


def timed(launches, reps=4):
    for l in launches[:2]:
        l.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for l in launches:
            l.run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * len(launches))


def main():
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)

    def w(*shape):
        return (torch.randn(shape, generator=g, device=dev) * 0.02).to(dt)
    n = 10
    Wg, Wu = [w(I, H) for _ in range(n)], [w(I, H) for _ in range(n)]
    Wd = [w(H, I) for _ in range(n)]
    Wo = [w(H, H) for _ in range(3 * n)]
    tunings = [("stream-K, 256-element chunks (default)", 0), ("128-element chunks", 128), ("copy-only", 64), ("4 stages", 4 << 8), ("register", 16)]
    for M in (1, 16, 32, 64):
        x, xi = w(M, H), w(M, I)
        act, y = torch.empty((M, I), dtype=dt, device=dev), torch.empty((M, H), dtype=dt, device=dev)
        for name, tuning in tunings:
            row = []
            ls = [DC.SkinnyLaunch([dict(A0=x, B0=Wg[c], C=act)], tuning) for c in range(n)]
            ms = timed(ls)
            row.append(f"up[11008x4096] {ls[0].bytes / ms / 1e6:6.0f}")
            ls = [DC.SkinnyLaunch([dict(A0=x, B0=Wg[c], B0u=Wu[c], C=act, epilogue=DC.SK_SILU_MUL)], tuning) for c in range(n)]
            ms = timed(ls)
            row.append(f"gate_up dual {ls[0].bytes / ms / 1e6:6.0f}")
            ls = [DC.SkinnyLaunch([dict(A0=xi, B0=Wd[c], C=y)], tuning) for c in range(n)]
            ms = timed(ls)
            row.append(f"down[4096x11008] {ls[0].bytes / ms / 1e6:6.0f}")
            ls = [DC.SkinnyLaunch([dict(A0=x, B0=Wo[c], C=y)], tuning) for c in range(3 * n)]
            ms = timed(ls)
            row.append(f"o[4096x4096] {ls[0].bytes / ms / 1e6:6.0f} ({ms * 1e3:.1f} us)")
            print(f"M={M:2d} {name:26s} | " + " | ".join(row) + "  GB/s", flush=True)


if __name__ == "__main__":
    main()
