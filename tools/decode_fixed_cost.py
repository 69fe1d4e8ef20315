"""What does one skinny-linear launch cost beyond its bytes?  Times launches whose span is a single iteration per CTA (no split row
blocks), launches with split row blocks, and a trivial kernel, back to back on a stream and inside a CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import _cabi  # noqa: E402
from modelcompose_b200 import decode as DC  # noqa: E402

dt = torch.bfloat16
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)


def w(*shape):
    return (torch.randn(shape, generator=g, device=dev) * 0.02).to(dt)


def timed(fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    stream_us = a.elapsed_time(b) * 1e3 / reps
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(50):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(4):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return stream_us, a.elapsed_time(b) * 1e3 / 200


M = 32
sms = torch.cuda.get_device_properties(dev).multi_processor_count
cases = [("1 iteration / CTA  (N=64*SMs, K=256)", 64 * sms, 256), ("2 iterations / CTA (N=64*SMs, K=512)", 64 * sms, 512),
         ("8 iterations / CTA (N=64*SMs, K=2048)", 64 * sms, 2048), ("split row blocks  (N=4096, K=4096: o_proj)", 4096, 4096),
         ("split row blocks  (N=1152, K=4096: LoRA down of q/k/v)", 1152, 4096)]
for name, N, K in cases:
    x, W, y = w(M, K), w(N, K), torch.empty((M, N), dtype=dt, device=dev)
    for tuning, tn in ((0, "stream-K"), (16, "register")):
        L = DC.SkinnyLaunch([dict(A0=x, B0=W, C=y)], tuning)
        s_us, g_us = timed(L.run)
        print(f"{name:58s} {tn:9s} {N * K * 2 / 1e6:7.1f} MB  stream {s_us:6.1f} us  graph {g_us:6.1f} us  (L2-resident weights)", flush=True)
# a trivial kernel for the launch gap alone
xr, wr, out = w(M, 4096), w(1, 4096).view(-1), torch.empty((M, 4096), dtype=dt, device=dev)
lib = _cabi.lib()


def rms():
    _cabi.check(lib.mc_rmsnorm(xr.data_ptr(), wr.data_ptr(), out.data_ptr(), M, 4096, 4096, 4096, 1e-5, _cabi.dtype_code(dt), _cabi.current_stream_ptr()), "rms")


s_us, g_us = timed(rms)
print(f"{'rmsnorm [32 x 4096] (launch gap reference)':58s} {'':9s} {0.26:7.1f} MB  stream {s_us:6.1f} us  graph {g_us:6.1f} us")
for pdl in (1,):
    prev = lib.mc_set_launch_mode(pdl)
    s_us, g_us = timed(rms)
    lib.mc_set_launch_mode(prev)
    print(f"{'rmsnorm, programmatic dependent launch':58s} {'':9s} {0.26:7.1f} MB  stream {s_us:6.1f} us  graph {g_us:6.1f} us")
