"""Development timing of the decode-step kernels at vicuna-7B shapes (CUDA events, L2 flushed by rotating over weight copies).

python tools/decode_dev.py [--m 1,8,32,64] [--tunings 0,16,32]
Prints GB/s of weight bytes per launch (roofline: the measured copy bandwidth in MEASURED_PEAKS.json)."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import _cabi  # noqa: E402
from modelcompose_b200 import decode as DC  # noqa: E402

H, I, V, R0 = 4096, 11008, 32000, 384
dt = torch.bfloat16


def timed(fn, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", default="1,8,32,64")
    ap.add_argument("--tunings", default="0,16,32")
    ap.add_argument("--copies", type=int, default=12)
    args = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)

    def w(*shape):
        return (torch.randn(shape, generator=g, device=dev) * 0.02).to(dt)
    # rotating weight sets: > L2 between two uses of the same matrix
    n = args.copies
    Wq = [[w(H, H) for _ in range(3)] for _ in range(n)]
    Wo = [w(H, H) for _ in range(n)]
    Wg, Wu = [w(I, H) for _ in range(n)], [w(I, H) for _ in range(n)]
    Wd = [w(H, I) for _ in range(n)]
    Wl = [w(V, H) for _ in range(max(2, n // 4))]
    Bq = [[w(H, 768) for _ in range(3)] for _ in range(n)]
    Aq = [[w(768, H) for _ in range(3)] for _ in range(n)]
    for M in [int(x) for x in args.m.split(",")]:
        x, xi = w(M, H), w(M, I)
        t = [w(M, R0) for _ in range(3)]
        cs = torch.ones(R0, dtype=torch.float32, device=dev)
        for tuning in [int(x) for x in args.tunings.split(",")]:
            rows = []

            def case(name, launches):
                ms = timed(lambda i: launches[i % len(launches)].run(), 4 * len(launches))
                rows.append(f"{name} {launches[0].bytes / ms / 1e6:7.0f} GB/s ({ms * 1e3:6.1f} us)")
            q, k, v, o, act = (torch.empty((M, H), dtype=dt, device=dev) for _ in range(4)), None, None, None, torch.empty((M, I), dtype=dt, device=dev)
            q = list(q)
            case("qkv", [DC.SkinnyLaunch([dict(A0=x, B0=Wq[c][i], C=q[i]) for i in range(3)], tuning) for c in range(n)])
            case("qkv+lora", [DC.SkinnyLaunch([dict(A0=x, B0=Wq[c][i], A1=t[i], B1=Bq[c][i][:, :R0], C=q[i]) for i in range(3)], tuning) for c in range(n)])
            case("down_qkv", [DC.SkinnyLaunch([dict(A0=x, B0=Aq[c][i][:R0], C=t[i], col_scale=cs, epilogue=DC.SK_COLSCALE) for i in range(3)], tuning) for c in range(n)])
            case("o+res", [DC.SkinnyLaunch([dict(A0=x, B0=Wo[c], C=q[3], residual=q[3], epilogue=DC.SK_RESIDUAL)], tuning) for c in range(n)])
            case("gate_up", [DC.SkinnyLaunch([dict(A0=x, B0=Wg[c], B0u=Wu[c], C=act, epilogue=DC.SK_SILU_MUL)], tuning) for c in range(n)])
            case("down+res", [DC.SkinnyLaunch([dict(A0=xi, B0=Wd[c], C=q[3], residual=q[3], epilogue=DC.SK_RESIDUAL)], tuning) for c in range(n)])
            lg = torch.empty((M, V), dtype=dt, device=dev)
            case("lm_head", [DC.SkinnyLaunch([dict(A0=x, B0=Wl[c], C=lg)], tuning) for c in range(len(Wl))])
            print(f"M={M:3d} tuning={tuning:3d} | " + " | ".join(rows), flush=True)
    # ---- attention
    lib = _cabi.lib()
    nH, D = 32, 128
    for B, L in ((32, 1000), (8, 3050), (16, 3600), (1, 4000)):
        cap = L + 8
        n_layers = max(2, int(2e9 // (B * cap * nH * D * 4)))
        kc = [torch.randn((B, nH, cap, D), device=dev, dtype=dt) for _ in range(n_layers)]
        vc = [torch.randn((B, nH, cap, D), device=dev, dtype=dt) for _ in range(n_layers)]
        qq = torch.randn((B, nH * D), device=dev, dtype=dt)
        out = torch.empty_like(qq)
        pos = torch.tensor([L - 1], dtype=torch.int32, device=dev)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        for splits in sorted({1, max(1, min(16, -(-4 * sms // (B * nH)))), max(1, min(32, -(-8 * sms // (B * nH))))}):
            scratch = torch.zeros(B * nH * splits * (D + 2), dtype=torch.float32, device=dev)
            counters = torch.zeros(B * nH, dtype=torch.int32, device=dev)

            def run(i):
                c = i % n_layers
                _cabi.check(lib.mc_decode_attention(qq.data_ptr(), kc[c].data_ptr(), vc[c].data_ptr(), cap, pos.data_ptr(), None, 0, out.data_ptr(),
                                                    nH * D, nH * D, B, nH, D, 1.0 / math.sqrt(D), splits, scratch.data_ptr(), counters.data_ptr(),
                                                    _cabi.dtype_code(dt), _cabi.current_stream_ptr()), "att")
            ms = timed(run, 4 * n_layers)
            byt = 2 * B * L * nH * D * 2
            print(f"attention B={B} L={L} splits={splits}: {byt / ms / 1e6:7.0f} GB/s ({ms * 1e3:6.1f} us)", flush=True)
        del kc, vc


if __name__ == "__main__":
    main()
