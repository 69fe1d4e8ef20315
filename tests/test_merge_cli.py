"""Merge CLI host logic (A1-A5) against the fixtures produced by the reference CLI, and against the live
reference when /root/reference is present."""
import json
import os

import pytest
import torch

from modelcompose_b200 import merge as M
from modelcompose_b200 import synthetic as syn

STRATEGY_C1 = "online-merge-reset-default-vision=0.5,default-audio=0.5"


def _write_inputs(golden, tmp_path):
    g = golden("merge_c1.pt")
    dirs = []
    for i, modal in enumerate(("vision", "audio")):
        sd, cfg = g["inputs"][modal]
        d = str(tmp_path / f"{modal}_ckpt")
        syn.save_checkpoint_dir(d, sd, cfg)
        dirs.append(d)
    return g, dirs


def _assert_same_outputs(out_dir, run, dirs):
    sd = torch.load(os.path.join(out_dir, "adapter_model.bin"), map_location="cpu")
    assert list(sd.keys()) == list(run["state_dict"].keys())
    for k in sd:
        assert sd[k].dtype == run["state_dict"][k].dtype and sd[k].shape == run["state_dict"][k].shape
        assert torch.equal(sd[k].view(torch.int16), run["state_dict"][k].view(torch.int16)), k
    assert open(os.path.join(out_dir, "config.json")).read() == run["config_json_text"]
    info = open(os.path.join(out_dir, "merge_info.txt")).read()
    assert info.replace(dirs[0], "{IN0}").replace(dirs[1], "{IN1}").replace(out_dir, "{OUT}") == run["merge_info"]


def test_online_merge_reset_cli_matches_reference_fixture(golden, tmp_path):
    g, dirs = _write_inputs(golden, tmp_path)
    out = str(tmp_path / "out-multimodal")
    M.main(dirs + ["-o", out, "--strategy", STRATEGY_C1])
    _assert_same_outputs(out, g["runs"][STRATEGY_C1], dirs)
    cfg = json.load(open(os.path.join(out, "config.json")))
    assert cfg["reset_scaling_weights"] == "default-vision=0.5,default-audio=0.5"
    assert cfg["vision_lora_r"] == 8 and cfg["audio_lora_alpha"] == 16


def test_synthetic_generator_reproduces_fixture_inputs(golden):
    g = golden("merge_c1.pt")
    for modal, seed, feat in (("vision", 100, 64), ("audio", 101, 48)):
        sd, cfg = syn.make_unimodal_checkpoint(modal, seed=seed, feat_dim=feat)
        ref_sd, ref_cfg = g["inputs"][modal]
        assert cfg == ref_cfg and list(sd) == list(ref_sd)
        assert all(torch.equal(sd[k].view(torch.int16), ref_sd[k].view(torch.int16)) for k in sd)


def test_error_behaviour(golden, tmp_path):
    g, dirs = _write_inputs(golden, tmp_path)
    with pytest.raises(UnboundLocalError):
        M.merge_checkpoints(dirs, str(tmp_path / "o"), "no-such-strategy")
    if not torch.cuda.is_available():  # the ties-* arithmetic has no CPU path: it must fail loudly, not fall back
        from modelcompose_b200 import _cabi
        with pytest.raises(_cabi.McError):
            M.merge_checkpoints(dirs, str(tmp_path / "o"), "ties-mean")
    with pytest.raises(AssertionError):
        M.get_modal_from_config({"lora_r": 8})
    # shared key without 'default' in its name → bare assert, as the reference (:101)
    sd, cfg = g["inputs"]["vision"]
    bad = dict(sd)
    bad["model.shared.weight"] = torch.zeros(2, dtype=torch.bfloat16)
    for i in range(2):
        syn.save_checkpoint_dir(str(tmp_path / f"bad{i}"), bad, cfg)
    with pytest.raises(AssertionError):
        M.merge_checkpoints([str(tmp_path / "bad0"), str(tmp_path / "bad1")], str(tmp_path / "o"), STRATEGY_C1)


def test_mm_projector_bin_fallback(golden, tmp_path):
    g = golden("merge_c1.pt")
    sd, cfg = g["inputs"]["vision"]
    d = tmp_path / "proj_only"
    d.mkdir()
    torch.save({k: v for k, v in sd.items() if "modal_projectors" in k}, d / "mm_projector.bin")
    json.dump(cfg, open(d / "config.json", "w"))
    merged, mcfg = M.merge_checkpoints([str(d)], str(tmp_path / "o-multimodal"), STRATEGY_C1)
    assert list(merged) == [k for k in sd if "modal_projectors" in k]


def test_live_reference_cli_agrees(golden, tmp_path):
    import _reference_loader as R
    if not R.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    g, dirs = _write_inputs(golden, tmp_path)
    ours, ref = str(tmp_path / "ours-multimodal"), str(tmp_path / "ref-multimodal")
    M.main(dirs + ["-o", ours, "--strategy", STRATEGY_C1])
    R.run_merge_cli(dirs + ["-o", ref, "--strategy", STRATEGY_C1])
    a = torch.load(os.path.join(ours, "adapter_model.bin"))
    b = torch.load(os.path.join(ref, "adapter_model.bin"))
    assert list(a) == list(b) and all(torch.equal(a[k].view(torch.int16), b[k].view(torch.int16)) for k in a)
    assert json.load(open(os.path.join(ours, "config.json"))) == json.load(open(os.path.join(ref, "config.json")))


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["sum", "mean"])
def test_sum_mean_cli_bit_exact_on_gpu(golden, tmp_path, strategy):
    g, dirs = _write_inputs(golden, tmp_path)
    out = str(tmp_path / f"out-{strategy}")
    M.main(dirs + ["-o", out, "--strategy", strategy])
    _assert_same_outputs(out, g["runs"][strategy], dirs)
