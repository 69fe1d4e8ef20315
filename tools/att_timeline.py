#!/usr/bin/env python
"""Development aid: clock64 timeline of one CTA of the attention kernel (tuning bit 9), printed as per-step deltas."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modelcompose_b200 import _cabi, linear as LN  # noqa: E402

B, S, nH = 8, 3046, 32
tuning = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0x13
q, k, v = (torch.randn((B * S, nH * 128), device="cuda", dtype=torch.bfloat16) for _ in range(3))
out = torch.empty_like(q)
for _ in range(3):
    LN.attention_causal(q, k, v, out, B, S, nH, 1 / math.sqrt(128), tuning=tuning)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    LN.attention_causal(q, k, v, out, B, S, nH, 1 / math.sqrt(128), tuning=tuning & ~0x200)
b.record()
torch.cuda.synchronize()
print(f"tuning {tuning:#x}: {a.elapsed_time(b) / 10:.3f} ms per launch")
LN.attention_causal(q, k, v, out, B, S, nH, 1 / math.sqrt(128), tuning=tuning | 0x200)
torch.cuda.synchronize()
raw = np.zeros(4 * 64 * 8 + 16, dtype=np.int64)
_cabi.check(_cabi.lib().mc_attention_debug_read(raw.ctypes.data_as(C.c_void_p), raw.nbytes), "debug_read")
buf = raw[:4 * 64 * 8].reshape(4, 64, 8)
life = raw[4 * 64 * 8:].reshape(2, 8)
for i, nm in enumerate(("heaviest query block", "lightest query block")):
    e = life[i]
    print(f"CTA life cycle, {nm}: entry->setup done {e[1] - e[0]}, ->first S ready {e[2] - e[1]}, ->softmax A loop end {e[3] - e[2]}, "
          f"softmax B loop end at +{e[4] - e[2]}, ->epilogue stored {e[5] - max(e[3], e[4])}, ->exit {e[6] - e[5]}; total {e[6] - e[0]}")
t0 = buf[buf > 0].min()
names = ["softmax A", "softmax B", "QK warp", "PV warp"]
print("softmax events: 0 top, 1 S ready, 2 ld done + S released, 3 max done, 4 exp done, 5 P buffer free, 6 st done, 7 arrived")
print("MMA warps: 0 top, 1 operand stage ready, 2 tile A issued, 3 tile B issued, 4 stage released")
for w in range(4):
    print(f"== {names[w]}: absolute start of step, then deltas between consecutive events")
    for j in range(8, 24):
        e = buf[w, j]
        n = 8 if w < 2 else 5
        if e[0] == 0 or buf[w, j + 1, 0] == 0:
            continue
        print(f"  step {j:2d} @ {e[0] - t0:7d}: " + " ".join(f"{e[i + 1] - e[i]:5d}" for i in range(n - 1)) +
              f" | step total {buf[w, j + 1, 0] - e[0] if buf[w, j + 1, 0] else 0}")
