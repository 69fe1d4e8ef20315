"""Host-side prompt / collator mirrors against the live reference when it is present, and on their own."""
import pytest
import torch

from modelcompose_b200 import data as D


class ToyTokenizer:
    """whitespace tokenizer with a BOS, enough to exercise tokenizer_modal_token"""
    bos_token_id = 1
    pad_token_id = 0

    def __call__(self, text):
        class R:
            pass
        r = R()
        r.input_ids = [1] + [3 + (sum(map(ord, w)) % 500) for w in text.split()]
        return r


PROMPTS = ["image: <image>\naudio: <audio>\nvideo: <video>\npoint: <point>\nwhat is this?", "no modality here",
           "<audio><image> back to back", "", "tail <video>", "<text> and <relrep> too"]


def test_constants_match_reference_values():
    assert D.MODAL_TOKEN_INDEXES == {"vision": -200, "relrep": -201, "text": -202, "audio": -203, "video": -204, "point": -205}
    assert D.MODAL_TOKEN_MAPPING["<image>"] == -200 and D.MODAL_TOKEN_MAPPING["<point>"] == -205


def test_tokenizer_modal_token():
    tok = ToyTokenizer()
    ids = D.tokenizer_modal_token(PROMPTS[0], tok)
    assert ids[0] == 1 and ids.count(1) == 1
    assert [t for t in ids if t < 0] == [-200, -203, -204, -205]
    assert D.tokenizer_modal_token("<audio><image> x", tok, return_tensors="pt").tolist()[:3] == [1, -203, -200]
    with pytest.raises(ValueError):
        D.tokenizer_modal_token("x", tok, return_tensors="np")


def test_live_reference_agrees():
    import _reference_loader as R
    if not R.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    import importlib.util
    import os
    import sys
    R._install_shells()
    spec = importlib.util.spec_from_file_location("ref_constants", os.path.join(R.REFERENCE_ROOT, "modelcompose", "constants.py"))
    const = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(const)
    assert const.MODAL_TOKEN_MAPPING == D.MODAL_TOKEN_MAPPING and const.MODAL_TOKEN_INDEXES == D.MODAL_TOKEN_INDEXES
    # mm_utils imports PIL / transformers symbols at module level; its two functions only need the constants
    src = open(os.path.join(R.REFERENCE_ROOT, "modelcompose", "mm_utils.py")).read()
    start, end = src.index("def split_string_by_list"), src.index("def get_model_name_from_path")
    ns = {"MODAL_TOKEN_MAPPING": const.MODAL_TOKEN_MAPPING, "torch": torch}
    exec(src[start:end], ns)
    tok = ToyTokenizer()
    for p in PROMPTS:
        assert ns["split_string_by_list"](p, list(const.MODAL_TOKEN_MAPPING)) == D.split_string_by_list(p, list(D.MODAL_TOKEN_MAPPING))
        assert ns["tokenizer_modal_token"](p, tok) == D.tokenizer_modal_token(p, tok)


def test_collator_and_buckets():
    col = D.FeatureCollator(pad_token_id=0, model_max_length=12)
    inst = [{"input_ids": [1, 5, -200, 7], "labels": [-100, 5, -100, 7], "modal_inputs": {"vision": [torch.ones(3, 4)]}},
            {"input_ids": [1, -203, 9, 9, 9, 2], "labels": [-100] * 6, "modal_inputs": {"audio": [torch.zeros(2, 5)], "vision": [torch.full((3, 4), 2.0)]}}]
    b = col(inst)
    assert b["input_ids"].shape == (2, 6) and b["input_ids"][0, 4:].tolist() == [0, 0]
    assert b["attention_mask"].tolist() == [[True] * 4 + [False] * 2, [True] * 6]
    assert b["labels"][0, 4:].tolist() == [-100, -100]
    assert b["modal_inputs"]["vision"].shape == (2, 3, 4) and b["modal_inputs"]["vision"][1, 0, 0] == 2.0
    assert b["modal_inputs"]["audio"].shape == (1, 2, 5)
    rows = {"vision": 3 + 10, "audio": 2 + 10}
    assert D.spliced_length(inst[0]["input_ids"], rows) == 3 + 13 and D.spliced_length(inst[1]["input_ids"], rows) == 5 + 12
    many = [{"input_ids": [1, -200] + [4] * n} for n in (3, 5, 3, 3, 5, 7, 3)]
    groups = list(D.bucket_by_length(many, 2, rows))
    assert sorted(map(sorted, groups)) == [[0, 2], [1, 4], [3, 6], [5]]
    for grp in groups:
        assert len({len(many[i]["input_ids"]) for i in grp}) == 1
    staged = D.PinnedBatch().load(b)
    out = staged.to("cpu")
    assert torch.equal(out["input_ids"], b["input_ids"]) and torch.equal(out["modal_inputs"]["audio"], b["modal_inputs"]["audio"])
