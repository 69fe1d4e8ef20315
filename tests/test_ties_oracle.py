"""Pin oracle/ties_oracle.py to tests/golden/ties.pt (outputs of the reference's own ties_merging.py and merge CLI) and,
when /root/reference is present, to the live reference on fresh random cases.  Bar: bit-exact, dtype and key order included."""
import contextlib
import copy
import hashlib
import io
import json
import os
import sys

import pytest
import torch

from modelcompose_b200 import merge as M
from modelcompose_b200 import synthetic as syn
from oracle import merge_oracle as MO
from oracle import ties_oracle as TO


def tensor_digest(t) -> str:
    return f"{t.dtype}|{tuple(t.shape)}|" + hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def test_kth_rank_matches_reference_expression():
    assert TO.kth_rank(1000, 20) == 800 and TO.kth_rank(1000, 0.3) == 700 and TO.kth_rank(7, 50) == 4
    assert TO.kth_rank(10, 1) == 10           # K = 1 is "1 percent": int(10 * 0.01) = 0 -> k = d
    for d, K in ((1000, 20), (12345, 0.7), (3, 99), (6738415616, 20)):
        assert M.ties_kth_rank(d, K) == TO.kth_rank(d, K)
    with pytest.raises(RuntimeError):
        TO.kth_rank(10, 100)                  # keeps everything -> kthvalue(0), the reference raises
    with pytest.raises(RuntimeError):
        M.ties_kth_rank(10, 100)


def test_do_merging_matches_reference_fixture(golden):
    g = golden("ties.pt")
    assert len(g["vectors"]) >= 9
    for case in g["vectors"]:
        for f, want in case["outputs"].items():
            got = TO.do_merging(case["checks"], K=case["K"], merge_func=f)
            assert list(got) == list(want), case["name"]
            for k in got:
                assert tensor_digest(got[k]) == tensor_digest(want[k]), (case["name"], f, k)


def test_fixture_covers_the_hard_cases(golden):
    """negative majority, exact cancellation and -0 outputs must all occur in the fixture, or it pins nothing"""
    g = golden("ties.pt")
    seen = {"neg_majority": False, "ambiguous": False, "neg_zero": False, "f32_mean": False}
    for case in g["vectors"]:
        flat = torch.vstack([TO.state_dict_to_vector(c) for c in case["checks"]])
        _, st = TO.ties_merge_flat(flat, case["K"], "sum")
        seen["neg_majority"] |= st["majority"] < 0
        seen["ambiguous"] |= st["ambiguous"] > 0
        mx = torch.cat([v.reshape(-1) for v in case["outputs"]["dis-max"].values()]).float()
        seen["neg_zero"] |= bool(((mx == 0) & torch.signbit(mx)).any())
        seen["f32_mean"] |= all(v.dtype == torch.float32 for v in case["outputs"]["dis-mean"].values())
    assert all(seen.values()), seen


def _cli_inputs(g):
    damc, same = syn.ties_cli_checkpoints()
    for fam, ck in (("damc", damc), ("same", same)):
        for m in ck:
            assert {k: tensor_digest(v) for k, v in ck[m][0].items()} == g["cli"]["inputs"][fam][m], "synthetic generator drifted"
    return {"damc": damc, "same": same}


def test_cli_strategies_match_reference_fixture(golden):
    g = golden("ties.pt")
    inputs = _cli_inputs(g)
    assert len(g["cli"]["runs"]) >= 7
    for name, run in g["cli"]["runs"].items():
        fam, strategy, K = name.split(":")
        ck = inputs[fam]
        sds = [ck[m][0] for m in ("vision", "audio")]
        cfgs = [copy.deepcopy(ck[m][1]) for m in ("vision", "audio")]
        merged, after, wtm = TO.merge_weights_extended(sds, cfgs, strategy, K=int(K), get_modal=MO.get_modal_from_config)
        if merged is None:  # the convert- prefix in front of an online-merge / sum strategy
            modal_names = [MO.get_modal_from_config(c) for c in cfgs]
            merged = {}
            for key, lst in wtm.items():
                if after.startswith("online-merge-"):
                    if len(lst) == 1:
                        merged[key] = lst[0]
                    else:
                        for mn, w in zip(modal_names, lst):
                            merged[key.replace("default", f"default-{mn}")] = w
                else:
                    merged[key] = MO.ref_sum(lst)
        assert list(merged) == list(run["digest"]), name
        for k in merged:
            assert tensor_digest(merged[k]) == run["digest"][k], (name, k)
        mcfg, after2 = MO.merge_configs(cfgs, after)
        assert json.dumps(mcfg, indent=4) == run["config_json_text"], name
        assert MO.merge_info_text(["{IN0}", "{IN1}"], after2, "{OUT}") == run["merge_info"], name


def test_live_reference_agrees_on_random_cases():
    import _reference_loader as R
    if not R.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    R._install_shells()
    sys.path.insert(0, os.path.join(R.REFERENCE_ROOT, "scripts", "model_composition"))
    try:
        import ties_merging as T
        g = torch.Generator().manual_seed(99)
        n_checked = 0
        for trial in range(60):
            dt = [torch.bfloat16, torch.float16, torch.float32][trial % 3]
            n_src, kind = 1 + trial % 4, trial % 4
            def gen(shape):
                if kind == 0:
                    return (torch.randn(shape, generator=g) * 0.02).to(dt)
                if kind == 1:
                    return torch.randint(-3, 4, shape, generator=g).to(dt)
                if kind == 2:
                    return (torch.randn(shape, generator=g) * 0.02 - 0.03).to(dt)
                return (torch.randn(shape, generator=g) * torch.randn(shape, generator=g) * 1e-3).to(dt)
            checks = [{"b": gen((int(torch.randint(1, 300, (1,), generator=g)),)), "a": gen((17, 5))} for _ in range(n_src)]
            checks = [{k: c[k][:checks[0][k].shape[0]] if k == "b" else c[k] for k in c} for c in checks]
            if len({c["b"].shape for c in checks}) != 1:
                n = min(c["b"].shape[0] for c in checks)
                checks = [dict(c, b=c["b"][:n]) for c in checks]
            K = [20, 50, 0.3, 1, 99, 5][trial % 6]
            for f in ("dis-sum", "dis-mean", "dis-max"):
                with contextlib.redirect_stdout(io.StringIO()):
                    want = T.do_merging(checks, K=K, merge_func=f)
                got = TO.do_merging(checks, K=K, merge_func=f)
                assert list(got) == list(want)
                for k in got:
                    assert tensor_digest(got[k]) == tensor_digest(want[k]), (trial, f, k)
                n_checked += 1
        assert n_checked == 180
    finally:
        sys.path.pop(0)
        sys.modules.pop("ties_merging", None)


METRIC_RTOL = 2e-5  # the reference reduces in fp32 (tree sums), the oracle and the kernel in fp64: stated tolerance


def test_interference_metrics_match_reference_fixture(golden):
    g = golden("ties.pt")
    n = 0
    for case in g["vectors"]:
        if "metrics" not in case:
            continue
        flat = torch.vstack([TO.state_dict_to_vector(c) for c in case["checks"]])
        got = TO.interference_metrics(flat, 50)
        for k, want in case["metrics"].items():
            assert abs(got[k] - want) <= METRIC_RTOL * max(1.0, abs(want)), (case["name"], k, got[k], want)
        n += 1
    assert n >= 8


def test_product_host_logic_without_gpu(tmp_path):
    """convert_delta_to_ft and the strategy plumbing are host code; the arithmetic has no CPU path and must say so."""
    a = {"x.default": torch.zeros(3), "u": torch.ones(2)}
    b = {"x.default": torch.ones(3)}
    ft, uniq = M.convert_delta_to_ft({"x.default": [a["x.default"], b["x.default"]], "u": [a["u"]]})
    assert [list(c) for c in ft] == [["x.default"], ["x.default"]] and list(uniq) == ["u"]
    if not torch.cuda.is_available():
        from modelcompose_b200 import _cabi
        with pytest.raises(_cabi.McError, match="no CPU fallback"):
            M.do_merging(ft, K=20, merge_func="dis-mean")
