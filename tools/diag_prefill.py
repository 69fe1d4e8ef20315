#!/usr/bin/env python
"""Development aid: where does a device-resident prefill step spend host and device time (no sync between steps)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN  # noqa: E402
from modelcompose_b200 import splice as SP  # noqa: E402
from modelcompose_b200 import model as MD  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
layers = int(os.environ.get("LAYERS", "8"))
model, desc, batch, ids_h, mask_h, feats_h, flops = BN.build_prefill("c3", dev, 0, layers=layers)
ids_d, mask_d = ids_h.to(dev), mask_h.to(dev)
feats_d = {m: v.to(dev) for m, v in feats_h.items()}
for _ in range(3):
    model.forward(ids_d, mask_d, modal_inputs=feats_d)
torch.cuda.synchronize()

orig_splice = SP.splice
orig_prefill = MD.MultimodalLlamaForCausalLM.prefill
log = []
def splice_t(*a, **k):
    t = time.perf_counter(); r = orig_splice(*a, **k); log.append(("splice", time.perf_counter() - t)); return r
def prefill_t(self, *a, **k):
    t = time.perf_counter(); r = orig_prefill(self, *a, **k); log.append(("prefill_issue", time.perf_counter() - t)); return r
MD.SP.splice = splice_t
MD.MultimodalLlamaForCausalLM.prefill = prefill_t

for sync in (False, True):
    log.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(5):
        t = time.perf_counter()
        model.forward(ids_d, mask_d, modal_inputs=feats_d)
        log.append(("forward_host", time.perf_counter() - t))
        if sync:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    print(f"sync_each_step={sync}: device {e0.elapsed_time(e1) / 5:.1f} ms/step, wall {(time.perf_counter() - w0) / 5 * 1e3:.1f} ms/step")
    for name in ("splice", "prefill_issue", "forward_host"):
        vals = [v * 1e3 for n, v in log if n == name]
        print("   ", name, " ".join(f"{v:.1f}" for v in vals), "ms")
