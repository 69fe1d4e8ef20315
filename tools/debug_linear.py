#!/usr/bin/env python
"""Bring-up aid for the tcgen05 linear kernel: small cases with structured diagnostics (not a test, not a bench)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402


def report(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    bad = err > (2 ** -7) * ref.abs() + 2 ** -8 * scale
    print(f"{name}: max_err {err.max().item():.4g} scale {scale:.4g} bad {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        M, N = bad.shape
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows:", rows[:16].tolist(), "... total", len(rows), " bad cols:", cols[:16].tolist(), "... total", len(cols))
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        r, c = int(rows[0]), int(cols[0])
        print(f"   got[{r},{c}:{c+8}]", got[r, c:c + 8].tolist())
        print(f"   ref[{r},{c}:{c+8}]", ref[r, c:c + 8].tolist())
        # per 8x(32) block error map of the first 128x256 tile
        blk = bad[:128, :256].float()
        if blk.numel():
            bm = blk.reshape(-1, 8, blk.shape[1]).amax(1)
            print("   rows-of-8 with errors:", bm.any(1).nonzero().flatten().tolist()[:32])
    return not bad.any()


def case(M, N, K, dtype=torch.bfloat16, tuning=0, kmask=None, seed=0, name=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, generator=g, device="cuda", dtype=torch.float32).to(dtype)
    B = torch.randn(N, K, generator=g, device="cuda", dtype=torch.float32).to(dtype)
    if kmask is not None:  # keep only some K columns non-zero
        z = torch.zeros(K, device="cuda", dtype=dtype)
        z[kmask] = 1
        A = A * z
    Cm = torch.full((M, N), 7.0, device="cuda", dtype=dtype)
    plan = LN.LinearPlan([LN.Problem(A, B, Cm)], tuning=tuning)
    plan.run()
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    return report(name or f"M{M} N{N} K{K} {dtype} tuning{tuning}", Cm, ref)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    ok &= case(128, 256, 16, kmask=None, name="single MMA K=16 (K tail zero-filled)")
    ok &= case(128, 256, 64, name="one k-block")
    ok &= case(128, 256, 64, kmask=slice(16, 32), name="one k-block, only K[16:32] non-zero (descriptor K-advance)")
    ok &= case(128, 128, 64, tuning=1, name="BN=128 one k-block")
    ok &= case(128, 256, 512, name="8 k-blocks (pipeline wrap)")
    ok &= case(256, 512, 256, name="2x2 tiles")
    ok &= case(1000, 1000, 1000, name="ragged 1000^3")
    ok &= case(4096, 4096, 4096, name="4096^3")
    ok &= case(4096, 4096, 4096, dtype=torch.float16, name="4096^3 fp16")
    ok &= case(300, 136, 72, tuning=1, name="ragged small BN=128")
    if "--two-cta" in sys.argv:
        ok &= case(512, 256, 64, tuning=3, name="2-CTA one k-block")
        ok &= case(512, 256, 512, tuning=3, name="2-CTA 8 k-blocks")
        ok &= case(1024, 768, 256, tuning=3, name="2-CTA 2x3 tiles")
        ok &= case(1000, 1000, 1000, tuning=3, name="2-CTA ragged 1000^3")
        ok &= case(4096, 4096, 4096, tuning=3, name="2-CTA 4096^3")
        ok &= case(300, 264, 72, tuning=3, name="2-CTA ragged small")
    print("ALL OK" if ok else "FAILURES", flush=True)
    # quick timing
    for (M, N, K) in [(8192, 8192, 8192), (30720, 4096, 4096), (30720, 11008, 4096), (30720, 4096, 11008)]:
        A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        Cm = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for tuning in ((3, 2) if "--two-cta" in sys.argv else (2, 1)):
            plan = LN.LinearPlan([LN.Problem(A, B, Cm)], tuning=tuning)
            for _ in range(3):
                plan.run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                plan.run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print(f"timing M{M} N{N} K{K} tuning={tuning}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            torch.matmul(A, B.t(), out=Cm)
        t0.record()
        for _ in range(10):
            torch.matmul(A, B.t(), out=Cm)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"   cuBLAS (library reference): {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
