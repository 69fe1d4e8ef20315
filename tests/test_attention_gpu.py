"""GPU parity of the tcgen05 causal attention kernel (through the C ABI) against an fp32 evaluation of the reference's eager
attention (multimodal_llama.py:295-312) on the same 16-bit inputs.

Bar (floating point): |err| <= tol * max|ref| with tol = 2^-7 (bf16) / 2^-10 (fp16) — the probabilities are rounded to the
storage dtype before the PV product (as the reference rounds its softmax output), everything else accumulates in fp32."""
import math

import pytest
import torch

from modelcompose_b200 import linear as LN

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10}


def reference(q, k, v, B, S, nH, scale):
    """eager attention in fp32: softmax(QK^T * scale + causal) V per (sequence, head)"""
    D = 128
    qf, kf, vf = (t.float().view(B, S, nH, D).transpose(1, 2) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(2, 3)) * scale
    s = s + torch.full((S, S), float("-inf"), device=q.device).triu(1)
    return torch.matmul(torch.softmax(s, dim=-1), vf).transpose(1, 2).reshape(B * S, nH * D)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,S,nH", [(1, 1, 1), (2, 128, 2), (1, 130, 1), (3, 300, 4), (2, 980, 3), (1, 2049, 2)])
def test_causal_attention_vs_fp32_reference(dtype, B, S, nH):
    g = torch.Generator(device="cuda").manual_seed(S)
    H = nH * 128
    T = B * S
    # q / k / v as column slices of one wider buffer (row stride != H), large-ish logits to make the softmax peaky
    buf = torch.randn((T, 3 * H + 64), generator=g, device="cuda").mul_(1.5).to(dtype)
    q, k, v = buf[:, :H], buf[:, H:2 * H], buf[:, 2 * H:3 * H]
    out = torch.full((T, H), 7.0, dtype=dtype, device="cuda")
    scale = 1.0 / math.sqrt(128)
    LN.attention_causal(q, k, v, out, B, S, nH, scale)
    torch.cuda.synchronize()
    ref = reference(q, k, v, B, S, nH, scale)
    err = (out.float() - ref).abs().max().item()
    assert err <= TOL[dtype] * ref.abs().max().item(), (err, ref.abs().max().item())
    # scattered output rows (the modality-major buffer order): same values, permuted rows
    perm = torch.randperm(T, generator=g, device="cuda").to(torch.int32)
    out2 = torch.zeros((T, H), dtype=dtype, device="cuda")
    LN.attention_causal(q, k, v, out2, B, S, nH, scale, out_rowmap=perm)
    torch.cuda.synchronize()
    assert torch.equal(out2[perm.long()], out)


@pytest.mark.parametrize("tuning", [0x10, 0x20, 0x30, 0x40, 0x120])
def test_kernel_variants_agree_with_reference(tuning):
    """exp2 split between the MUFU pipe and the FMA-pipe polynomial (0 .. 3 pairs of 4), lazy vs eager rescaling: every
    variant meets the same bar; peaky logits (amplitude 6) exercise the rescale path"""
    B, S, nH = 2, 700, 2
    g = torch.Generator(device="cuda").manual_seed(tuning)
    H, T = nH * 128, B * S
    q, k, v = (torch.randn((T, H), generator=g, device="cuda").mul_(6.0 if i < 2 else 1.0).to(torch.bfloat16) for i in range(3))
    out = torch.empty((T, H), dtype=torch.bfloat16, device="cuda")
    scale = 1.0 / math.sqrt(128)
    LN.attention_causal(q, k, v, out, B, S, nH, scale, tuning=tuning)
    torch.cuda.synchronize()
    ref = reference(q, k, v, B, S, nH, scale)
    err = (out.float() - ref).abs().max().item()
    assert err <= TOL[torch.bfloat16] * ref.abs().max().item(), (hex(tuning), err, ref.abs().max().item())


def test_argument_errors():
    x = torch.zeros((4, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        LN.attention_causal(x, x, x, x.clone(), 1, 4, 1, 1.0)          # head_dim 64
    y = torch.zeros((4, 128), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        LN.attention_causal(y, y, y, y.clone(), 1, 4, 1, 1.0, out_rowmap=torch.zeros(4, dtype=torch.int64, device="cuda"))
