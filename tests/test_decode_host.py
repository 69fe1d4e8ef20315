"""Host-side pieces of the decode path that need no GPU: the key/value cache container (transformers-4.31 views, growth) and the
ctypes mirror of mc_skinny_desc_t against the header."""
import ctypes
import os
import re

import torch

from modelcompose_b200 import decode as DC
from modelcompose_b200 import model as MD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kv_cache_layout_growth_and_legacy_views():
    c = MD.KVCache(n_layers=2, B=3, capacity=5, n_heads=2, head_dim=4, dtype=torch.float32, device="cpu")
    assert len(c) == 2 and c.k[0].shape == (3, 2, 5, 4)  # [B, heads, capacity, head_dim]
    for l in range(2):
        c.k[l][:, :, :4] = torch.arange(3 * 2 * 4 * 4, dtype=torch.float32).view(3, 2, 4, 4) + 100 * l
        c.v[l][:, :, :4] = -c.k[l][:, :, :4]
    c.length = 4
    k0, v0 = c[0]
    assert k0.shape == (3, 2, 4, 4) and torch.equal(v0, -k0)  # past_key_value layout [B, heads, length, head_dim]
    before = [t[:, :, :4].clone() for t in c.k]
    c.grow(9)
    assert c.capacity == 9 and c.k[1].shape == (3, 2, 9, 4) and c.length == 4
    assert all(torch.equal(t[:, :, :4], b) for t, b in zip(c.k, before))
    assert len(c.legacy()) == 2 and c.legacy()[1][0].shape == (3, 2, 4, 4)


def test_skinny_desc_mirror_matches_header():
    text = open(os.path.join(ROOT, "include", "modelcompose_b200.h")).read()
    body = text[text.index("typedef struct mc_skinny_desc {"):text.index("} mc_skinny_desc_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if decl:
            for name in decl.replace("*", " ").split(",") if "," in decl and "int32_t" in decl else [decl.replace("*", " ")]:
                fields.append(name.split()[-1])
    assert fields == [f[0] for f in DC.SkinnyDesc._fields_], (fields, [f[0] for f in DC.SkinnyDesc._fields_])
    # pointers and int64 leading dimensions are 8 bytes, the four sizes and the epilogue int32
    assert ctypes.sizeof(DC.SkinnyDesc) == 4 * 4 + 8 * 15 + 8 + 8  # 16 + fifteen 8-byte members + col_scale + (epilogue + pad)


def test_decode_module_constants():
    assert DC.MAX_M == 64 and (DC.SK_NONE, DC.SK_RESIDUAL, DC.SK_COLSCALE, DC.SK_SILU_MUL) == (0, 1, 2, 3)
    text = open(os.path.join(ROOT, "include", "modelcompose_b200.h")).read()
    assert "#define MC_SKINNY_MAX_M 64" in text and "MC_SKINNY_EPI_SILU_MUL = 3" in text
