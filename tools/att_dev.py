#!/usr/bin/env python
"""Development harness of the causal attention kernel: parity of one kernel variant (mc_attention_causal_tuned `tuning`)
against an fp32 evaluation of the eager attention at awkward shapes, then CUDA-event timing at the prefill shapes.

    python tools/att_dev.py --tuning 0x32 [--no-parity] [--no-time] [--cudnn]

Run one variant per process under `timeout` (a variant that deadlocks must not take the rest of the sweep with it)."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import linear as LN  # noqa: E402

F = torch.nn.functional


def reference(q, k, v, B, S, nH, scale):
    D = 128
    qf, kf, vf = (t.float().view(B, S, nH, D).transpose(1, 2) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(2, 3)) * scale
    s = s + torch.full((S, S), float("-inf"), device=q.device).triu(1)
    return torch.matmul(torch.softmax(s, dim=-1), vf).transpose(1, 2).reshape(B * S, nH * D)


def parity(tuning):
    ok = True
    for dtype, tol in ((torch.bfloat16, 2.0 ** -7), (torch.float16, 2.0 ** -10)):
        for B, S, nH, amp in ((1, 1, 1, 1.5), (2, 128, 2, 1.5), (1, 130, 1, 1.5), (3, 300, 4, 1.5), (2, 980, 3, 1.5), (1, 2049, 2, 1.5),
                              (2, 700, 2, 6.0)):
            g = torch.Generator(device="cuda").manual_seed(S)
            H, T = nH * 128, B * S
            buf = torch.randn((T, 3 * H + 64), generator=g, device="cuda").mul_(amp).to(dtype)
            q, k, v = buf[:, :H], buf[:, H:2 * H], buf[:, 2 * H:3 * H]
            out = torch.full((T, H), 7.0, dtype=dtype, device="cuda")
            scale = 1.0 / math.sqrt(128)
            LN.attention_causal(q, k, v, out, B, S, nH, scale, tuning=tuning)
            torch.cuda.synchronize()
            ref = reference(q, k, v, B, S, nH, scale)
            d = (out.float() - ref).abs()
            err, mx = d.max().item(), ref.abs().max().item()
            good = err <= tol * mx and bool(torch.isfinite(out.float()).all())
            ok &= good
            extra = ""
            if not good:
                # where: per 128-row tile of the first sequence / per 32-column block of the first head
                rows = d.view(B, S, H)[0].max(dim=1).values
                per_tile = [round(rows[i:i + 128].max().item(), 3) for i in range(0, S, 128)]
                cols = d.view(B, S, H)[0, :, :128].max(dim=0).values
                per_col = [round(cols[i:i + 32].max().item(), 3) for i in range(0, 128, 32)]
                extra = f" row tiles {per_tile} col blocks {per_col}"
            print(f"  parity tuning={tuning:#x} {str(dtype)[6:]} B={B} S={S} nH={nH} amp={amp}: err {err:.5f} / tol {tol * mx:.5f} "
                  f"{'ok' if good else 'FAIL'}{extra}", flush=True)
            perm = torch.randperm(T, generator=g, device="cuda").to(torch.int32)
            out2 = torch.zeros((T, H), dtype=dtype, device="cuda")
            LN.attention_causal(q, k, v, out2, B, S, nH, scale, out_rowmap=perm, tuning=tuning)
            torch.cuda.synchronize()
            if not torch.equal(out2[perm.long()], out):
                ok = False
                print("  rowmap scatter FAIL", flush=True)
    return ok


def timing(tuning, cudnn):
    shapes = ((32, 980), (8, 3046), (16, 3569))
    if os.environ.get("ATT_SHAPES"):
        shapes = tuple(tuple(int(x) for x in s.split("x")) for s in os.environ["ATT_SHAPES"].split(","))
    for B, S in shapes:
        nH, D = 32, 128
        T, H = B * S, nH * D
        q, k, v = (torch.randn((T, H), device="cuda", dtype=torch.bfloat16) for _ in range(3))
        out = torch.empty_like(q)
        scale = 1.0 / math.sqrt(D)
        flops = 4.0 * B * nH * D * S * (S + 1) / 2

        def native():
            LN.attention_causal(q, k, v, out, B, S, nH, scale, tuning=tuning)

        def lib():
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
                return F.scaled_dot_product_attention(q.view(B, S, nH, D).transpose(1, 2), k.view(B, S, nH, D).transpose(1, 2),
                                                      v.view(B, S, nH, D).transpose(1, 2), is_causal=True, scale=scale)
        for name, fn in ((f"native {tuning:#x}", native),) + ((("cudnn", lib),) if cudnn else ()):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print(f"  B={B} S={S}: {name} {ms:.3f} ms ({flops / ms / 1e9:.0f} TF/s causal)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tuning", type=lambda x: int(x, 0), default=0)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--cudnn", action="store_true")
    a = ap.parse_args()
    print(f"== tuning {a.tuning:#x}", flush=True)
    good = True
    if not a.no_parity:
        good = parity(a.tuning)
        print(f"  parity {'ALL OK' if good else 'FAILED'}", flush=True)
    if not a.no_time and good:
        timing(a.tuning, a.cudnn)
