"""ncu target: a few launches of the decode-step kernels at vicuna-7B shapes, M = batch = 32 (weights rotated so they come from HBM).
ncu --set full -k regex:'streamk|decode_attention' -s 8 -c 8 python tools/profile_decode.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import _cabi  # noqa: E402
from modelcompose_b200 import decode as DC  # noqa: E402

H, I = 4096, 11008
dt = torch.bfloat16
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)


def w(*shape):
    return (torch.randn(shape, generator=g, device=dev) * 0.02).to(dt)


M = int(os.environ.get("MC_PROFILE_M", "32"))
n = 6
x = w(M, H)
act, y = torch.empty((M, I), dtype=dt, device=dev), [torch.empty((M, H), dtype=dt, device=dev) for _ in range(3)]
gu = [DC.SkinnyLaunch([dict(A0=x, B0=w(I, H), B0u=w(I, H), C=act, epilogue=DC.SK_SILU_MUL)]) for _ in range(n)]
qkv = [DC.SkinnyLaunch([dict(A0=x, B0=w(H, H), C=y[i]) for i in range(3)]) for _ in range(n)]
B, L, nH, D = 32, 1000, 32, 128
cap = L + 8
kc = [torch.randn((B, nH, cap, D), device=dev, dtype=dt) for _ in range(n)]
vc = [torch.randn((B, nH, cap, D), device=dev, dtype=dt) for _ in range(n)]
q = torch.randn((B, nH * D), device=dev, dtype=dt)
out = torch.empty_like(q)
pos = torch.tensor([L - 1], dtype=torch.int32, device=dev)
lib = _cabi.lib()
for i in range(n):  # launch order: gate_up, qkv, attention, repeated; the first two rounds are warm-up
    gu[i].run()
    qkv[i].run()
    _cabi.check(lib.mc_decode_attention(q.data_ptr(), kc[i].data_ptr(), vc[i].data_ptr(), cap, pos.data_ptr(), None, 0, out.data_ptr(), nH * D, nH * D,
                                        B, nH, D, 1.0 / math.sqrt(D), 1, None, None, _cabi.dtype_code(dt), _cabi.current_stream_ptr()), "att")
torch.cuda.synchronize()
print("done")
