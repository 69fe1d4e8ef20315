#!/usr/bin/env python
"""Generate tests/golden/*.pt by executing the UNMODIFIED reference (/root/reference).

Run in the authoring container only:  ``python tests/golden/make_golden.py``.
The reference has no tests or golden vectors of its own; these fixtures
are outputs of the reference's own code on seeded synthetic inputs,
so the oracle and the CUDA path are both pinned to what the reference actually computes.

Fixtures:
  merge_c1.pt    reference merge CLI on two tiny checkpoints (online-merge-reset, sum, mean)      A1-A5
  linear_c1.pt   reference LocalLoraLinear (+ scaling dict after reset coefficients)              A8, A9
  projector.pt   reference build_vision_projector mlp2x_gelu / linear                             A10
  splice.pt      reference prepare_inputs_labels_for_multimodal on 6 cases                        A11-A13
  prefill_c1.pt  reference decoder layers (routed attention + MLP) + thin wrapper → logits        A14-A16
  ties.pt        reference ties_merging.do_merging on seeded vectors (3 dtypes x 3 functions) and the      §8(f)3
                 reference CLI with ties-* / convert-* strategies on two 1-layer checkpoints
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _reference_loader as R  # noqa: E402
from modelcompose_b200 import synthetic as syn  # noqa: E402

STRATEGY_C1 = "online-merge-reset-default-vision=0.5,default-audio=0.5"


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()


def tensor_digest(t) -> str:
    """dtype, shape and sha256 of the raw bytes (bit-exact comparisons without storing the tensor)."""
    return f"{t.dtype}|{tuple(t.shape)}|" + hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def c1_checkpoints():
    v_sd, v_cfg = syn.make_unimodal_checkpoint("vision", seed=100, feat_dim=64)
    a_sd, a_cfg = syn.make_unimodal_checkpoint("audio", seed=101, feat_dim=48)
    return (v_sd, v_cfg), (a_sd, a_cfg)


def gen_merge():
    (v_sd, v_cfg), (a_sd, a_cfg) = c1_checkpoints()
    out = {"inputs": {"vision": (v_sd, v_cfg), "audio": (a_sd, a_cfg)}, "runs": {}}
    with tempfile.TemporaryDirectory() as tmp:
        vdir, adir = os.path.join(tmp, "vision_ckpt"), os.path.join(tmp, "audio_ckpt")
        syn.save_checkpoint_dir(vdir, v_sd, v_cfg)
        syn.save_checkpoint_dir(adir, a_sd, a_cfg)
        for strategy in (STRATEGY_C1, "sum", "mean"):
            odir = os.path.join(tmp, "out-multimodal-" + strategy.split("-")[0])
            R.run_merge_cli([vdir, adir, "-o", odir, "--strategy", strategy])
            sd = torch.load(os.path.join(odir, "adapter_model.bin"), map_location="cpu")
            cfg = json.load(open(os.path.join(odir, "config.json")))
            info = open(os.path.join(odir, "merge_info.txt")).read()
            info = info.replace(vdir, "{IN0}").replace(adir, "{IN1}").replace(odir, "{OUT}")
            out["runs"][strategy] = {"state_dict": sd, "config": cfg, "merge_info": info,
                                     "config_json_text": open(os.path.join(odir, "config.json")).read()}
    torch.save(out, os.path.join(HERE, "merge_c1.pt"))
    return out


def composed_cfg(merged_cfg: dict):
    kw = {k: v for k, v in merged_cfg.items() if k not in ("model_type", "architectures")}
    return R.make_config(**kw)


def gen_linear(merge):
    llama = R.load_llama()
    run = merge["runs"][STRATEGY_C1]
    cfg = composed_cfg(run["config"])
    modal_names = R.infer_modals_restated(cfg)
    base = syn.make_base_llm(seed=1)
    out = {"modal_names": modal_names, "cases": {}}
    for lname in ("self_attn.q_proj", "mlp.down_proj"):
        o_f, i_f = syn.linear_shape(syn.TINY, lname)
        lin = llama.LocalLoraLinear(modal_names, i_f, o_f, cfg.lora_r, cfg.lora_alpha, cfg.lora_dropout, bias=False,
                                    reset_scaling_weights=cfg.reset_scaling_weights).eval()
        prefix = f"model.layers.0.{lname}."
        sd = {k[len(prefix):]: v for k, v in run["state_dict"].items() if k.startswith(prefix)}
        sd["weight"] = base[prefix + "weight"]
        missing, unexpected = lin.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
        # the un-merged "default" adapter has no weights in a merged checkpoint (it stays at its init)
        assert not unexpected, unexpected
        g = torch.Generator().manual_seed(7)
        x = torch.randn(2, 16, i_f, generator=g)
        case = {"scaling": dict(lin.scaling), "adapters": list(lin.lora_A.keys()),
                "default_adapter_names": list(lin.default_adapter_names),
                "merge_default_weights": lin.merge_default_weights, "x": x, "out": {}}
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            lin_dt = lin.to(dt)
            with torch.no_grad():
                y = lin_dt(x.to(dt), active_adapters=modal_names)
                y0 = lin_dt(x.to(dt))
            case["out"][str(dt)] = {k: v.clone() for k, v in y.items()}
            case["out"][str(dt)]["__base__"] = y0.clone()
            lin = lin.float()
        out["cases"][lname] = case
    torch.save(out, os.path.join(HERE, "linear_c1.pt"))


def gen_projector(merge):
    proj = R.load_projector_builder()
    out = {}
    run = merge["runs"][STRATEGY_C1]

    class C:  # config namespace as the builder reads it (getattr)
        hidden_size = 256
    for modal, feat, builder, tkey, hkey in (("vision", 64, proj.build_vision_projector, "mm_projector_type", "mm_hidden_size"),
                                             ("audio", 48, proj.build_audio_projector, "mm_audio_projector_type", "mm_audio_hidden_size")):
        c = C()
        setattr(c, tkey, "mlp2x_gelu")
        setattr(c, hkey, feat)
        m = builder(c).eval()
        prefix = f"model.modal_projectors.{modal}."
        m.load_state_dict({k[len(prefix):]: v.float() for k, v in run["state_dict"].items() if k.startswith(prefix)})
        g = torch.Generator().manual_seed(2000 + feat)
        x = torch.randn(3, 11, feat, generator=g)
        res = {"x": x}
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            with torch.no_grad():
                res[str(dt)] = m.to(dt)(x.to(dt)).clone()
            m = m.float()
        out[modal] = res
    # 'linear' projector type
    c = C()
    c.mm_projector_type = "linear"
    c.mm_hidden_size = 64
    torch.manual_seed(5)
    m = proj.build_vision_projector(c).eval()
    x = torch.randn(2, 7, 64)
    with torch.no_grad():
        out["linear"] = {"x": x, "weight": m.weight.detach().clone(), "bias": m.bias.detach().clone(),
                         str(torch.float32): m(x).clone()}
    torch.save(out, os.path.join(HERE, "projector.pt"))


def _splice_case(name, cfg_kw, ids, attn, labels, feats, modal_inputs_keys, prefix=None, suffix=None, vocab=50, H=16,
                 seed=0):
    """Run the real prepare_inputs_labels_for_multimodal with identity projectors and fake encoders."""
    import torch.nn as nn
    cfg = R.make_config(hidden_size=H, intermediate_size=32, num_attention_heads=2, num_key_value_heads=2,
                        num_hidden_layers=1, vocab_size=vocab, **cfg_kw)
    g = torch.Generator().manual_seed(seed)
    emb = nn.Embedding(vocab, H)
    with torch.no_grad():
        emb.weight.copy_(torch.randn(vocab, H, generator=g))
    modals = [m for m in R.infer_modals_restated(cfg) if m != "default"]
    host = R.make_splice_host(cfg, emb, {m: nn.Identity() for m in modals})
    modal_inputs = {}
    for m in modal_inputs_keys:
        modal_inputs[m] = {"audio_inputs": feats[m], "audio_padding_mask": None} if m == "audio" else feats[m]
    # modalities configured but absent from modal_inputs take encoder.dummy_inputs in the reference;
    # our fake encoders have none, so tests always pass every configured modality unless stated.
    rec = {"name": name, "cfg_kw": cfg_kw, "input_ids": ids, "attention_mask": attn, "labels": labels,
           "features": {m: (v.reshape(v.shape[0], -1, v.shape[-1]) if v.dim() == 4 else v) for m, v in feats.items()},
           "modal_inputs_keys": list(modal_inputs_keys), "embed": emb.weight.detach().clone(),
           "prefix": prefix, "suffix": suffix, "modals": modals}
    try:
        with torch.no_grad():
            r = host.prepare_inputs_labels_for_multimodal(ids, attn, None, labels, modal_inputs, prefix, suffix)
        rec["out"] = {"attention_mask": r[1], "inputs_embeds": r[3], "labels": r[4], "modal_attention_mask": r[5]}
    except Exception as e:  # the reference's own failure modes are part of the contract
        rec["raises"] = type(e).__name__
    return rec


def gen_splice():
    V, A, P, VID = -200, -203, -205, -204
    H = 16
    g = torch.Generator().manual_seed(11)

    def f(n_blocks, n_tok):
        return torch.randn(n_blocks, n_tok, H, generator=g)

    def ids_rows(rows):
        return torch.tensor(rows, dtype=torch.int64)

    cases = []
    va = dict(mm_vision_tower="v", mm_audio_encoder="a")
    # 1. equal-length, vision+audio, bool mask, no labels (the inference shape)
    ids = ids_rows([[1, 5, 6, V, 7, A, 8, 9], [1, 9, 8, V, 3, A, 4, 2]])
    cases.append(_splice_case("equal_bool", va, ids, torch.ones_like(ids, dtype=torch.bool), None,
                              {"audio": f(2, 3), "vision": f(2, 4)}, ["audio", "vision"]))
    # 2. int64 attention mask (what HF generate passes) → masks come out int64
    cases.append(_splice_case("equal_int64", va, ids, torch.ones_like(ids), None,
                              {"audio": f(2, 3), "vision": f(2, 4)}, ["audio", "vision"], seed=1))
    # 3. labels + ragged (sample 1 has no audio block) → right padding, left-extended attention mask
    ids = ids_rows([[1, 5, 6, V, 7, A, 8, 9], [1, 9, 8, V, 3, 4, 4, 2]])
    labels = ids.clone()
    labels[labels < 0] = -100
    cases.append(_splice_case("ragged_labels", va, ids, torch.ones_like(ids, dtype=torch.bool), labels,
                              {"audio": f(2, 3), "vision": f(2, 4)}, ["audio", "vision"], seed=2))
    # 4. ragged without labels → the reference raises UnboundLocalError
    cases.append(_splice_case("ragged_nolabels", va, ids, torch.ones_like(ids, dtype=torch.bool), None,
                              {"audio": f(2, 3), "vision": f(2, 4)}, ["audio", "vision"], seed=3))
    # 5. two vision blocks in one sample, sentinel first/last, global cursor across the batch, prefix+suffix
    ids = ids_rows([[V, 5, 6, V], [V, 9, 8, V]])
    pre = {"default": torch.zeros(1, 2, H), "vision": f(1, 2)}
    suf = {"default": torch.zeros(1, 1, H), "vision": f(1, 1)}
    cases.append(_splice_case("double_block_prefix", dict(mm_vision_tower="v"), ids, torch.ones_like(ids, dtype=torch.bool),
                              None, {"vision": f(4, 3)}, ["vision"], prefix=pre, suffix=suf, seed=4))
    # 6. four modalities, video+vision+audio+point order, one sample without any sentinel (hacky path), labels
    vp = dict(mm_vision_tower="v", mm_audio_encoder="a", mm_video_encoder="vid", mm_point_encoder="p")
    ids = ids_rows([[1, VID, 2, V, 3, A, 4, P, 5, 6], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]])
    labels = ids.clone()
    labels[labels < 0] = -100
    cases.append(_splice_case("four_modal_nosentinel", vp, ids, torch.ones_like(ids, dtype=torch.bool), labels,
                              {"audio": f(1, 2), "vision": f(1, 3), "video": torch.randn(1, 2, 3, H, generator=g), "point": f(1, 2)},
                              ["audio", "vision", "video", "point"], seed=5))
    # 7. attention mask with left padding zeros, equal-length, int64
    ids = ids_rows([[0, 0, 1, V, 7, A, 8, 9], [1, 9, 8, V, 3, A, 4, 2]])
    attn = torch.tensor([[0, 0, 1, 1, 1, 1, 1, 1], [1] * 8])
    cases.append(_splice_case("leftpad_int64", va, ids, attn, None,
                              {"audio": f(2, 3), "vision": f(2, 4)}, ["audio", "vision"], seed=6))
    torch.save(cases, os.path.join(HERE, "splice.pt"))


def gen_prefill(merge):
    """Unmodified MultimodalLlamaDecoderLayer x2 + restated thin wrapper (SURVEY §8(c) last row)."""
    llama = R.load_llama()
    from transformers.models.llama.modeling_llama import LlamaRMSNorm
    run = merge["runs"][STRATEGY_C1]
    cfg = composed_cfg(run["config"])
    modal_names = R.infer_modals_restated(cfg)
    base = syn.make_base_llm(seed=1)
    torch.manual_seed(0)
    layers = [llama.MultimodalLlamaDecoderLayer(cfg).eval() for _ in range(cfg.num_hidden_layers)]
    # the reference loader resets every adapter (kaiming A, zero B) then loads the merged checkpoint;
    # the plain 'default' adapter is absent from a merged checkpoint → zero B → contributes nothing.
    for li, layer in enumerate(layers):
        sd = {}
        pre = f"model.layers.{li}."
        for k, v in run["state_dict"].items():
            if k.startswith(pre):
                sd[k[len(pre):]] = v.float()
        for k, v in base.items():
            if k.startswith(pre):
                sd[k[len(pre):]] = v.float()
        missing, unexpected = layer.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(".default.weight" in m for m in missing), missing
    norm = LlamaRMSNorm(cfg.hidden_size, eps=cfg.rms_norm_eps)
    with torch.no_grad():
        norm.weight.copy_(base["model.norm.weight"].float())
    # inputs: B=2, S'=24 tokens, segments text(4) audio(6) text(2) vision(8) text(4)
    g = torch.Generator().manual_seed(42)
    Bsz, S = 2, 24
    x = torch.randn(Bsz, S, cfg.hidden_size, generator=g) * 0.5
    seg = torch.tensor([0] * 4 + [1] * 6 + [0] * 2 + [2] * 8 + [0] * 4)
    masks = {"audio": (seg == 1)[None].expand(Bsz, S).clone(), "vision": (seg == 2)[None].expand(Bsz, S).clone()}
    masks["default"] = (torch.stack([masks[k] for k in masks]).sum(0) == 0)
    pos = torch.arange(S)[None]
    out = {"x": x, "masks": masks, "modal_names": modal_names, "base_digest": sd_digest(base), "out": {}}
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        add = torch.full((S, S), torch.finfo(dt).min, dtype=dt).triu(1)[None, None].expand(Bsz, 1, S, S)
        h = x.to(dt)
        hs = []
        with torch.no_grad():
            for layer in layers:
                layer.to(dt)
                h = layer(h, attention_mask=add, modal_attention_mask=masks, position_ids=pos)[0]
                hs.append(h.clone())
                layer.float()
            hn = norm.to(dt)(h)
            logits = torch.nn.functional.linear(hn, base["lm_head.weight"].to(dt))
            norm.float()
            # no-mask path (decode-style: default adapter only)
            h0 = x.to(dt)
            for layer in layers:
                layer.to(dt)
                h0 = layer(h0, attention_mask=add, modal_attention_mask=None, position_ids=pos)[0]
                layer.float()
        out["out"][str(dt)] = {"hidden": hs, "final_norm": hn.clone(), "logits": logits.clone(),
                               "hidden_nomask": h0.clone()}
    torch.save(out, os.path.join(HERE, "prefill_c1.pt"))


def ties_vectors():
    """Seeded do_merging inputs: gaussian, tie-heavy integer, negative-majority and exact-cancellation data."""
    cases = []
    g = torch.Generator().manual_seed(4242)
    for i, (dt, n_src, kind, K) in enumerate([
            (torch.bfloat16, 3, "gauss", 20), (torch.bfloat16, 2, "ints", 50), (torch.bfloat16, 4, "neg", 20),
            (torch.float16, 3, "gauss", 0.3), (torch.float16, 2, "ints", 20), (torch.float32, 3, "gauss", 20),
            (torch.float32, 2, "ints", 70), (torch.float32, 3, "neg", 5), (torch.bfloat16, 1, "gauss", 20)]):
        def gen(shape):
            if kind == "gauss":
                return (torch.randn(shape, generator=g) * 0.02).to(dt)
            if kind == "ints":
                return torch.randint(-3, 4, shape, generator=g).to(dt)
            return (torch.randn(shape, generator=g) * 0.02 - 0.03).to(dt)
        checks = [{"w.b": gen((37, 29)), "w.a": gen((1500,)), "w.c": gen((3,))} for _ in range(n_src)]
        cases.append({"checks": checks, "K": K, "name": f"{i}-{str(dt)[6:]}-{n_src}src-{kind}"})
    return cases


def gen_ties():
    import contextlib
    import io
    R._install_shells()
    sys.path.insert(0, os.path.join(R.REFERENCE_ROOT, "scripts", "model_composition"))
    import ties_merging as T
    out = {"vectors": [], "cli": {"inputs": {}, "runs": {}}}
    for case in ties_vectors():
        res = {}
        for f in ("dis-sum", "dis-mean", "dis-max"):
            with contextlib.redirect_stdout(io.StringIO()):
                res[f] = dict(T.do_merging(case["checks"], K=case["K"], merge_func=f))
        out["vectors"].append(dict(case, outputs=res))
    import calculate_metrics as CM
    for entry in out["vectors"]:
        if len(entry["checks"]) < 2:
            continue
        flat = torch.vstack([T.state_dict_to_vector({k: v.float() for k, v in c.items()}, []) for c in entry["checks"]])
        trunc, *_ = T.topk_values_mask(flat.clone(), K=50, return_mask=False)
        entry["metrics"] = {"L2": float(CM.L2(flat)), "Cosine": float(CM.cos_sim(flat)),
                            "SSD": float(CM.soft_sign_dissimilarity(flat)), "TSSD": float(CM.soft_sign_dissimilarity(trunc))}
    sys.path.pop(0)
    sys.modules.pop("ties_merging", None)
    # CLI: ties-* on two 1-layer DAMC checkpoints, convert-* on two 1-layer `same` checkpoints
    damc, same = syn.ties_cli_checkpoints()
    # inputs are regenerated from their seeds by the tests; only their digests are stored
    out["cli"]["inputs"] = {fam: {m: {k: tensor_digest(v) for k, v in ck[m][0].items()} for m in ck}
                            for fam, ck in (("damc", damc), ("same", same))}
    runs = [("damc", "ties-mean", 20), ("damc", "ties-sum", 35), ("damc", "ties-max", 20),
            ("same", "convert-drop-mean", 20), ("same", "convert-ties-sum", 20),
            ("same", "convert-online-merge-reset-default-vision=0.5,default-audio=0.5", 20), ("same", "convert-sum", 20)]
    with tempfile.TemporaryDirectory() as tmp:
        dirs = {}
        for fam, ck in (("damc", damc), ("same", same)):
            dirs[fam] = []
            for m in ("vision", "audio"):
                d = os.path.join(tmp, f"{fam}_{m}")
                syn.save_checkpoint_dir(d, ck[m][0], ck[m][1])
                dirs[fam].append(d)
        for i, (fam, strategy, K) in enumerate(runs):
            odir = os.path.join(tmp, f"out-multimodal-{i}")
            with contextlib.redirect_stdout(io.StringIO()):
                R.run_merge_cli(dirs[fam] + ["-o", odir, "--strategy", strategy, "-K", str(K)])
            info = open(os.path.join(odir, "merge_info.txt")).read()
            info = info.replace(dirs[fam][0], "{IN0}").replace(dirs[fam][1], "{IN1}").replace(odir, "{OUT}")
            sd = torch.load(os.path.join(odir, "adapter_model.bin"), map_location="cpu")
            # per-key digests for every run; the full tensors only for two runs (ties outputs are views of one vector: clone)
            run = {"digest": {k: tensor_digest(v) for k, v in sd.items()},
                   "config_json_text": open(os.path.join(odir, "config.json")).read(), "merge_info": info}
            if strategy in ("ties-mean", "convert-drop-mean"):
                run["state_dict"] = {k: v.clone() for k, v in sd.items()}
            if strategy == "ties-mean":  # the reference's calculate_metrics.py on this merged directory (paths are still valid)
                sys.path.insert(0, os.path.join(R.REFERENCE_ROOT, "scripts", "model_composition"))
                import calculate_metrics as CM2
                with contextlib.redirect_stdout(io.StringIO()):
                    CM2.calculate_metrics(odir)
                sys.path.pop(0)
                for mname in ("calculate_metrics", "ties_merging"):
                    sys.modules.pop(mname, None)
                run["merge_metrics_txt"] = open(os.path.join(odir, "merge_metrics.txt")).read()
            out["cli"]["runs"][f"{fam}:{strategy}:{K}"] = run
    torch.save(out, os.path.join(HERE, "ties.pt"))
    return out


def main():
    assert R.reference_available(), "needs /root/reference (authoring container)"
    torch.set_num_threads(1)  # deterministic CPU reductions
    if "--only-ties" in sys.argv:
        gen_ties()
        return
    merge = gen_merge()
    gen_linear(merge)
    gen_projector(merge)
    gen_splice()
    gen_prefill(merge)
    gen_ties()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
