"""GPU parity of the splice kernels (through the C ABI) against oracle/splice_oracle.py and the fixtures produced by
the reference's own prepare_inputs_labels_for_multimodal (tests/golden/splice.pt).  Bar: bit-exact."""
import pytest
import torch

from modelcompose_b200 import _cabi
from modelcompose_b200 import splice as SP
from oracle import splice_oracle as SO

pytestmark = pytest.mark.gpu


def cuda(x):
    if x is None:
        return None
    if isinstance(x, dict):
        return {k: cuda(v) for k, v in x.items()}
    return x.cuda().contiguous()


def run_case(case, dtype=torch.float32):
    feats = {m: case["features"][m].to(dtype) for m in case["modals"]}
    pre = {k: v.to(dtype) for k, v in case["prefix"].items()} if case["prefix"] else None
    suf = {k: v.to(dtype) for k, v in case["suffix"].items()} if case["suffix"] else None
    return SP.prepare_inputs_labels_for_multimodal(
        cuda(case["input_ids"]), cuda(case["attention_mask"]), None, cuda(case["labels"]), cuda(feats), cuda(pre),
        cuda(suf), cuda(case["embed"].to(dtype)), case["modal_inputs_keys"])


def assert_same(got, want_attn, want_embeds, want_labels, want_masks, name):
    none_ids, attn, pkv, embeds, labels, masks = got
    assert none_ids is None and pkv is None
    assert embeds.dtype == want_embeds.dtype and torch.equal(embeds.cpu(), want_embeds), name
    assert attn.dtype == want_attn.dtype and torch.equal(attn.cpu(), want_attn), name
    if want_labels is None:
        assert labels is None
    else:
        assert torch.equal(labels.cpu(), want_labels), name
    assert list(masks.keys()) == list(want_masks.keys()), name
    for k in masks:
        assert masks[k].dtype == want_masks[k].dtype, (name, k)
        assert torch.equal(masks[k].cpu(), want_masks[k]), (name, k)


def test_reference_fixtures_bit_exact(golden):
    for case in golden("splice.pt"):
        if "raises" in case:
            with pytest.raises(Exception) as ei:
                run_case(case)
            assert type(ei.value).__name__ == case["raises"], case["name"]
            continue
        ref = case["out"]
        assert_same(run_case(case), ref["attention_mask"], ref["inputs_embeds"], ref["labels"],
                    ref["modal_attention_mask"], case["name"])


def random_case(seed, B, S, H, vocab, modal_rows, dtype, equal=True, with_labels=False, mask_dtype=torch.bool,
                prefix=0, suffix=0):
    """Random ids with sentinels; equal=True gives every sample the same sentinel multiset (inference shape)."""
    g = torch.Generator().manual_seed(seed)
    names = list(modal_rows)
    ids = torch.randint(3, vocab, (B, S), generator=g)
    counts = {m: 0 for m in names}
    for b in range(B):
        chosen = names if equal else [m for m in names if torch.rand(1, generator=g).item() < 0.6]
        pos = torch.randperm(S, generator=g)[:len(chosen) + 1].tolist()
        for m, p in zip(chosen, pos):
            ids[b, p] = SO.MODAL_TOKEN_INDEXES[m]
            counts[m] += 1
        if not equal and b % 3 == 0 and chosen:  # a second block of the first modality
            ids[b, pos[-1]] = SO.MODAL_TOKEN_INDEXES[chosen[0]]
            counts[chosen[0]] += 1
    feats = {m: torch.randn(max(counts[m], 1), modal_rows[m], H, generator=g).to(dtype) for m in names}
    pre = {m: torch.randn(1, prefix, H, generator=g).to(dtype) for m in names} if prefix else None
    suf = {m: torch.randn(1, suffix, H, generator=g).to(dtype) for m in names} if suffix else None
    embed = torch.randn(vocab, H, generator=g).to(dtype)
    attn = (torch.rand(B, S, generator=g) > 0.1).to(mask_dtype)
    labels = None
    if with_labels:
        labels = ids.clone()
        labels[labels < 0] = -100
    return ids, attn, labels, embed, feats, pre, suf


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("equal,with_labels,mask_dtype", [(True, False, torch.bool), (True, False, torch.int64),
                                                          (False, True, torch.bool), (False, True, torch.int64),
                                                          (True, True, torch.bool)])
def test_random_batches_vs_oracle(dtype, equal, with_labels, mask_dtype):
    rows = {"audio": 7, "vision": 13, "video": 33, "point": 5}
    for seed, (B, S, H, pfx, sfx) in enumerate([(4, 12, 64, 0, 0), (9, 300, 256, 5, 5), (3, 40, 8, 2, 0), (1, 6, 4096, 0, 3)]):
        if H * torch.empty((), dtype=dtype).element_size() % 16:
            continue
        ids, attn, labels, embed, feats, pre, suf = random_case(100 + seed, B, S, H, 500, rows, dtype, equal, with_labels,
                                                                mask_dtype, pfx, sfx)
        want = SO.splice(ids, attn, labels, embed, SO.add_prefix_suffix(feats, pre, suf))
        got = SP.prepare_inputs_labels_for_multimodal(cuda(ids), cuda(attn), None, cuda(labels), cuda(feats), cuda(pre),
                                                      cuda(suf), cuda(embed))
        assert_same(got, want[0], want[1], want[2], want[3], (seed, dtype, equal))


def test_modal_id_and_lengths():
    rows = {"audio": 3, "vision": 4}
    ids, attn, labels, embed, feats, pre, suf = random_case(7, 5, 20, 32, 100, rows, torch.bfloat16, equal=False,
                                                            with_labels=True, prefix=1, suffix=2)
    r = SP.splice(cuda(ids), cuda(attn), cuda(labels), cuda(embed), cuda(feats), cuda(pre), cuda(suf))
    want = SO.splice(ids, attn, labels, embed, SO.add_prefix_suffix(feats, pre, suf))
    mid = r.modal_id.cpu()
    for i, m in enumerate(r.modal_names):
        assert torch.equal(mid == i + 1, want[3][m].bool())
    assert torch.equal(mid == 0, want[3]["default"])
    assert max(r.out_len) == r.inputs_embeds.shape[1]
    assert r.algorithmic_bytes == 5 * r.inputs_embeds.shape[1] * 2 * 32 * 2


def test_no_sentinel_batch_matches_hacky_path():
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 50, (3, 9), generator=g)
    embed = torch.randn(50, 16, generator=g)
    feats = {"audio": torch.randn(1, 2, 16, generator=g), "vision": torch.randn(1, 3, 16, generator=g)}
    attn = torch.ones(3, 9, dtype=torch.bool)
    want = SO.splice(ids, attn, None, embed, feats)
    got = SP.prepare_inputs_labels_for_multimodal(cuda(ids), cuda(attn), None, None, cuda(feats), None, None, cuda(embed))
    assert_same(got, want[0], want[1], want[2], want[3], "hacky")


def test_error_paths():
    g = torch.Generator().manual_seed(2)
    embed = torch.randn(50, 16, generator=g).cuda()
    feats = {"vision": torch.randn(1, 3, 16, generator=g).cuda()}
    attn = torch.ones(2, 4, dtype=torch.bool).cuda()
    # more sentinels than feature blocks (reference: IndexError)
    ids = torch.tensor([[1, -200, 2, 3], [1, -200, 2, 3]]).cuda()
    with pytest.raises(_cabi.McError, match="feature blocks"):
        SP.splice(ids, attn, None, embed, feats)
    # sentinel of a modality that was not configured / id beyond the vocabulary
    for bad in (-203, 50):
        ids = torch.tensor([[1, bad, 2, 3], [1, 4, 2, 3]]).cuda()
        with pytest.raises(_cabi.McError, match="sample 0"):
            SP.splice(ids, attn, None, embed, feats)
    with pytest.raises(ValueError):
        SP.splice(ids.cpu(), attn, None, embed, feats)
    # decode step early return (multimodal_arch.py:290-293)
    one = torch.tensor([[5], [6]]).cuda()
    out = SP.prepare_inputs_labels_for_multimodal(one, attn, None, None, None, None, None, embed)
    assert out[0] is one and out[3] is None and out[5] is None


def test_full_size_c3_properties():
    """BASELINE config 3 shape (batch 32, 576 image + 256 audio + 128 text, H=4096, 5+5 prefix/suffix): the oracle
    would take minutes, so check size-independent properties: every output row is bit-identical to its source row
    (row checksums over a permutation-invariant re-gather), masks partition the rows, lengths add up."""
    B, H, V = 32, 4096, 32000
    g = torch.Generator(device="cuda").manual_seed(3)
    text = 128
    S = text + 2
    ids = torch.randint(3, V, (B, S), generator=g, device="cuda")
    ids[:, 40] = -200
    ids[:, 45] = -203
    embed = torch.randn(V, H, generator=g, device="cuda", dtype=torch.bfloat16)
    feats = {"audio": torch.randn(B, 256, H, generator=g, device="cuda", dtype=torch.bfloat16),
             "vision": torch.randn(B, 576, H, generator=g, device="cuda", dtype=torch.bfloat16)}
    pre = {m: torch.randn(1, 5, H, generator=g, device="cuda", dtype=torch.bfloat16) for m in feats}
    suf = {m: torch.randn(1, 5, H, generator=g, device="cuda", dtype=torch.bfloat16) for m in feats}
    attn = torch.ones(B, S, dtype=torch.int64, device="cuda")
    r = SP.splice(ids, attn, None, embed, feats, pre, suf)
    Sp = text + 576 + 256 + 20
    assert r.inputs_embeds.shape == (B, Sp, H) and r.out_len == [Sp] * B
    e = r.inputs_embeds
    # layout: ids[:40] | pre_v, vision, suf_v | ids[41:45] | pre_a, audio, suf_a | ids[46:]
    assert torch.equal(e[:, :40], embed[ids[:, :40]])
    o = 40
    assert torch.equal(e[:, o:o + 5], pre["vision"].expand(B, -1, -1))
    assert torch.equal(e[:, o + 5:o + 581], feats["vision"])
    assert torch.equal(e[:, o + 581:o + 586], suf["vision"].expand(B, -1, -1))
    o += 586
    assert torch.equal(e[:, o:o + 4], embed[ids[:, 41:45]])
    o += 4
    assert torch.equal(e[:, o:o + 5], pre["audio"].expand(B, -1, -1))
    assert torch.equal(e[:, o + 5:o + 261], feats["audio"])
    assert torch.equal(e[:, o + 261:o + 266], suf["audio"].expand(B, -1, -1))
    o += 266
    assert torch.equal(e[:, o:], embed[ids[:, 46:]])
    m = r.modal_attention_mask
    assert (m["audio"] + m["vision"] + m["default"].to(torch.int64) == 1).all()
    assert m["vision"].sum().item() == B * 586 and m["audio"].sum().item() == B * 266
    assert r.attention_mask.shape == (B, Sp) and r.attention_mask.all()
