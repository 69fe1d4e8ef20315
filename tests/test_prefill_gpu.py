"""GPU parity of the composed-model prefill (splice -> projectors -> routed decoder layers -> logits).

Compared against (a) the fixtures the UNMODIFIED reference decoder layers produced (tests/golden/prefill_c1.pt) and
(b) the CPU oracle run on the same inputs.  Tolerance (bf16 tolerance stated per north_star; max-abs and cosine are
printed): logits / hidden max-abs <= 2^-5 (bf16) or 2^-8 (fp16) of the tensor's max magnitude and cosine >= 0.9995 /
0.99999 — the reference rounds after every op in 16 bits, the kernels accumulate in fp32 and round once per linear."""
import json
import os

import pytest
import torch

from modelcompose_b200 import builder as BD
from modelcompose_b200 import model as MD
from modelcompose_b200 import synthetic as syn
from oracle import merge_oracle as MO
from oracle import model_oracle as XO
from oracle import splice_oracle as SO

pytestmark = pytest.mark.gpu

STRATEGY_C1 = "online-merge-reset-default-vision=0.5,default-audio=0.5"
DTYPES = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}
MAXABS = {"torch.bfloat16": 2 ** -5, "torch.float16": 2 ** -8}
COS = {"torch.bfloat16": 0.9995, "torch.float16": 0.99999}


def compare(got, ref, key, what):
    got, ref = got.float().cpu().flatten(), ref.float().cpu().flatten()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    print(f"{what} [{key}]: max-abs {err:.4g} (scale {scale:.4g}, ratio {err / scale:.3g}) cosine {cos:.7f}")
    assert err <= MAXABS[key] * scale and cos >= COS[key], (what, err, scale, cos)


def tiny_model(golden, dtype):
    run = golden("merge_c1.pt")["runs"][STRATEGY_C1]
    cfg = MD.MultimodalConfig.from_dict(run["config"])
    base = syn.make_base_llm(seed=1)
    return MD.MultimodalLlamaForCausalLM(cfg, base, run["state_dict"], device="cuda", dtype=dtype), run, base


@pytest.mark.parametrize("key", list(DTYPES))
def test_decoder_layers_vs_reference_fixture(golden, key):
    dtype = DTYPES[key]
    model, run, base = tiny_model(golden, dtype)
    assert model.modal_names == ["default", "audio", "vision"]
    assert model.default_adapter_names == ["default-audio", "default-vision"]
    g = golden("prefill_c1.pt")
    ref = g["out"][key]
    x = g["x"].to(dtype).cuda()
    mid = torch.zeros(x.shape[:2], dtype=torch.uint8)
    for i, m in enumerate(model.modal_names):
        mid[g["masks"][m]] = i
    logits, _, hidden = model.prefill(x, mid.cuda(), None, output_hidden_states=True)
    torch.cuda.synchronize()
    compare(hidden[1], ref["hidden"][0], key, "hidden after layer 0")
    compare(hidden[2], ref["final_norm"], key, "final norm")
    compare(logits, ref["logits"], key, "logits")
    ws = next(iter(model._ws.values()))
    compare(ws.x.view_as(x), ref["hidden"][1], key, "hidden after layer 1")
    # modal_id=None: every token takes the default adapter (decode-style path, multimodal_llama.py:436-438,:703-704)
    model.prefill(x, None, None)
    compare(ws.x.view_as(x), ref["hidden_nomask"], key, "hidden (no modality mask)")


@pytest.mark.parametrize("key", list(DTYPES))
def test_end_to_end_forward_vs_oracle(golden, key):
    """input_ids with sentinels + encoder features -> projector -> splice -> routed layers -> logits, vs the CPU oracle
    evaluated in the same 16-bit dtype."""
    dtype = DTYPES[key]
    model, run, base = tiny_model(golden, dtype)
    sd = run["state_dict"]
    g = torch.Generator().manual_seed(9)
    B, n_text = 3, 20
    ids = syn.make_prompt_ids(B, ["vision", "audio"], n_text, 1000, seed=3, modal_token_indexes=SO.MODAL_TOKEN_INDEXES, n_head=6)
    feats = {"audio": torch.randn(B, 9, 48, generator=g).to(dtype), "vision": torch.randn(B, 14, 64, generator=g).to(dtype)}
    attn = torch.ones_like(ids)
    out = model.forward(ids.cuda(), attn.cuda(), modal_inputs={k: v.cuda() for k, v in feats.items()})
    torch.cuda.synchronize()
    # ---- oracle
    proj = {}
    for m in ("audio", "vision"):
        pre = f"model.modal_projectors.{m}."
        proj[m] = XO.projector_forward(feats[m], [sd[pre + "0.weight"].to(dtype), sd[pre + "2.weight"].to(dtype)],
                                       [sd[pre + "0.bias"].to(dtype), sd[pre + "2.bias"].to(dtype)])
    pre_t = {m: sd[f"prefix_tokens.{m}"].to(dtype) for m in proj}
    suf_t = {m: sd[f"suffix_tokens.{m}"].to(dtype) for m in proj}
    am, embeds, _, masks = SO.splice(ids, attn, None, base["model.embed_tokens.weight"].to(dtype),
                                     SO.add_prefix_suffix(proj, pre_t, suf_t))
    names, scaling, dnames = MO.effective_scaling(["default", "audio", "vision"], 8, 16, run["config"]["reset_scaling_weights"])
    layers = []
    for li in range(2):
        layer = {"input_layernorm": base[f"model.layers.{li}.input_layernorm.weight"].to(dtype),
                 "post_attention_layernorm": base[f"model.layers.{li}.post_attention_layernorm.weight"].to(dtype)}
        for ln in syn.LINEAR_NAMES:
            p = f"model.layers.{li}.{ln}."
            A = {k[len(p) + 7:-7]: v.to(dtype) for k, v in sd.items() if k.startswith(p + "lora_A.")}
            Bm = {k[len(p) + 7:-7]: v.to(dtype) for k, v in sd.items() if k.startswith(p + "lora_B.")}
            layer[ln.split(".")[1]] = XO.LinearParams(base[p + "weight"].to(dtype), A, Bm, scaling, dnames)
        layers.append(layer)
    bmasks = {k: v.bool() for k, v in masks.items()}
    ordered = {m: bmasks[m] for m in ["default", "audio", "vision"]}
    logits, _ = XO.model_forward(embeds, layers, base["model.norm.weight"].to(dtype), base["lm_head.weight"].to(dtype),
                                 ordered, ["default", "audio", "vision"], 4, 1e-5)
    assert out.logits.shape == logits.shape
    assert torch.equal(out.modal_id.cpu() == 1, bmasks["audio"]) and torch.equal(out.modal_id.cpu() == 2, bmasks["vision"])
    compare(out.logits, logits, key, "end-to-end logits")


def test_loader_from_disk_and_text_only(tmp_path, golden):
    """merge CLI output dir + base dir -> load_pretrained_model (fp16 like the reference) -> forward."""
    from modelcompose_b200 import merge as MG
    (v_sd, v_cfg), (a_sd, a_cfg) = golden("merge_c1.pt")["inputs"]["vision"], golden("merge_c1.pt")["inputs"]["audio"]
    vdir, adir, odir, bdir = (str(tmp_path / n) for n in ("vision", "audio", "out-multimodal", "base"))
    syn.save_checkpoint_dir(vdir, v_sd, v_cfg)
    syn.save_checkpoint_dir(adir, a_sd, a_cfg)
    MG.merge_checkpoints([vdir, adir], odir, STRATEGY_C1)
    os.makedirs(bdir)
    torch.save(syn.make_base_llm(seed=1), os.path.join(bdir, "pytorch_model.bin"))
    tok, model, procs, ctx = BD.load_pretrained_model(odir, bdir, "out-multimodal")
    assert tok is None and procs is None and ctx == 2048 and model.dtype == torch.float16
    assert json.load(open(os.path.join(odir, "config.json")))["reset_scaling_weights"] == "default-vision=0.5,default-audio=0.5"
    ids = torch.randint(3, 1000, (2, 17), generator=torch.Generator().manual_seed(0)).cuda()
    out = model.forward(ids, torch.ones_like(ids))
    assert out.logits.shape == (2, 17, 1000) and torch.isfinite(out.logits).all()
    with pytest.raises(NotImplementedError):
        BD.load_pretrained_model(odir, bdir, "llava-thing")
    with pytest.raises(NotImplementedError):
        model.forward(ids[:, :1], torch.ones_like(ids), past_key_values=[()])
