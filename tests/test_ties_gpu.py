"""GPU parity of the TIES merge (through the C ABI) against oracle/ties_oracle.py, the reference fixtures
(tests/golden/ties.pt) and — at sizes the oracle cannot reach — size-independent properties.

Bar: bit-exact (raw words, signed zeros and result dtype included)."""
import copy
import hashlib
import os

import pytest
import torch

from modelcompose_b200 import merge as M
from modelcompose_b200 import synthetic as syn
from oracle import ties_oracle as TO

pytestmark = pytest.mark.gpu

INT_VIEW = {torch.float32: torch.int32, torch.float16: torch.int16, torch.bfloat16: torch.int16}
SIZES = [1, 7, 16, 17, 4096, 8192, 8193, 40000, (1 << 18) + 5]


def tensor_digest(t) -> str:
    t = t.cpu()
    return f"{t.dtype}|{tuple(t.shape)}|" + hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def bits_equal(a, b) -> bool:
    a, b = a.cpu(), b.cpu()
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a.view(INT_VIEW[a.dtype]), b.view(INT_VIEW[b.dtype]))


def make_sources(n_src, sizes, dtype, seed, kind):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_src):
        lst = []
        for n in sizes:
            if kind == "gauss":
                t = torch.randn(n, generator=g) * 0.02
            elif kind == "ints":      # ties at the threshold, exact cancellations
                t = torch.randint(-3, 4, (n,), generator=g).float()
            elif kind == "neg":       # negative majority sign
                t = torch.randn(n, generator=g) * 0.02 - 0.03
            else:                     # wide dynamic range incl. subnormals of the 16-bit types
                t = (torch.randn(n, generator=g) * torch.exp(torch.randn(n, generator=g) * 6.0) * 1e-4).clamp(-6e4, 6e4)  # finite in fp16
            lst.append(t.to(dtype))
        out.append(lst)
    return out


def oracle_merge(srcs, K, func):
    flat = torch.vstack([torch.cat(lst) for lst in srcs])
    merged, stats = TO.ties_merge_flat(flat, K, func)
    outs, pos = [], 0
    for t in srcs[0]:
        outs.append(merged[pos:pos + t.numel()])
        pos += t.numel()
    return outs, stats


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n_src", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("kind,K", [("gauss", 20), ("ints", 50), ("neg", 20), ("wide", 0.3), ("ints", 5), ("gauss", 99)])
def test_device_plan_bit_exact(dtype, n_src, kind, K):
    srcs = make_sources(n_src, SIZES, dtype, seed=17 * n_src + len(kind), kind=kind)
    dev = [[t.cuda() for t in lst] for lst in srcs]
    for func in ("sum", "mean", "max"):
        odt = torch.float32 if func == "mean" else dtype
        outs = [torch.full((n,), 7.0, dtype=odt, device="cuda") for n in SIZES]
        plan = M.TiesPlan(dev, outs)
        plan.run(K, func)
        st = plan.stats()
        want, ost = oracle_merge(srcs, K, func)
        assert st["thresholds"] == [float(x) for x in ost["thresholds"]], (func, st, ost)
        assert (st["n_pos"], st["n_neg"], st["majority"], st["n_ambiguous"]) == (ost["n_pos"], ost["n_neg"], int(ost["majority"]), ost["ambiguous"])
        assert st["n_pos"] + st["n_neg"] + st["n_zero"] + st["n_ambiguous"] == sum(SIZES)
        for t, n in enumerate(SIZES):
            assert bits_equal(outs[t], want[t]), (func, n, dtype, kind, int((outs[t].cpu().float() != want[t].float()).sum()))
        passes = 3 if dtype == torch.float32 else 1
        assert plan.algorithmic_bytes == sum(SIZES) * ((passes + 1) * n_src * srcs[0][0].element_size() + outs[0].element_size())
        plan.close()


@pytest.mark.parametrize("K", [90, 20])
def test_bf16_magnitudes_beyond_the_fp16_pattern_range(K):
    """The packed trim compares bf16 pairs as fp16 bit patterns, which is only valid below 2^121 (pattern 0x7c00): values beyond
    it are kept by the compare (right for any sane threshold), a THRESHOLD beyond it sends the kernel down its scalar path.
    Sources with every magnitude in [2^120, 2^124): keeping the top 90 % puts the threshold below 2^121 with data on both sides
    of it, keeping the top 20 % puts the threshold itself beyond."""
    g = torch.Generator().manual_seed(5)
    srcs = []
    for _ in range(3):
        lst = []
        for n in SIZES:
            mag = torch.exp2(120.0 + 4.0 * torch.rand(n, generator=g))
            lst.append((mag * (torch.randint(0, 2, (n,), generator=g).float() * 2 - 1)).to(torch.bfloat16))
        srcs.append(lst)
    dev = [[t.cuda() for t in lst] for lst in srcs]
    for func in ("sum", "mean", "max"):
        odt = torch.float32 if func == "mean" else torch.bfloat16
        outs = [torch.full((n,), 7.0, dtype=odt, device="cuda") for n in SIZES]
        plan = M.TiesPlan(dev, outs)
        plan.run(K, func)
        st = plan.stats()
        want, ost = oracle_merge(srcs, K, func)
        assert st["thresholds"] == [float(x) for x in ost["thresholds"]]
        assert (min(st["thresholds"]) >= 2.0 ** 121) == (K == 20) and (max(st["thresholds"]) < 2.0 ** 121) == (K == 90)
        assert (st["n_pos"], st["n_neg"], st["n_ambiguous"]) == (ost["n_pos"], ost["n_neg"], ost["ambiguous"])
        for t in range(len(SIZES)):
            assert bits_equal(outs[t], want[t]), (func, K, t)
        plan.close()


def test_fix_pass_only_when_needed():
    """positive majority: the speculative pass is final; negative majority with cancellations: the listed elements are
    recomputed (1); MAX with a negative majority or an overflowing list: dense re-merge (2) — all bit-exact vs the oracle"""
    g = torch.Generator().manual_seed(5)
    pos = [[(torch.randn(50000, generator=g) * 0.02 + 0.03).to(torch.bfloat16)] for _ in range(3)]
    out = [torch.empty(50000, dtype=torch.bfloat16, device="cuda")]
    plan = M.TiesPlan([[t.cuda() for t in lst] for lst in pos], out)
    plan.run(20, "sum")
    st = plan.stats()
    assert st["majority"] == 1 and st["fix_pass_ran"] == 0
    plan.run(20, "max")
    assert plan.stats()["fix_pass_ran"] == 0
    neg = [[torch.randint(-3, 3, (50000,), generator=g).to(torch.bfloat16)] for _ in range(2)]
    plan2 = M.TiesPlan([[t.cuda() for t in lst] for lst in neg], out)
    for func, want_fix in (("sum", 1), ("max", 2)):
        plan2.run(50, func)
        st = plan2.stats()
        assert st["majority"] == -1 and st["n_ambiguous"] > 0 and st["fix_pass_ran"] == want_fix
        assert bits_equal(out[0], oracle_merge(neg, 50, func)[0][0])
    # more majority-dependent elements than the list holds (2^20): dense fallback
    n = 9_000_000
    big = [[torch.randint(-2, 2, (n,), generator=g).to(torch.bfloat16)] for _ in range(2)]
    out_big = [torch.empty(n, dtype=torch.bfloat16, device="cuda")]
    plan3 = M.TiesPlan([[t.cuda() for t in lst] for lst in big], out_big)
    plan3.run(60, "sum")
    st = plan3.stats()
    assert st["majority"] == -1 and st["n_ambiguous"] > (1 << 20) and st["fix_pass_ran"] == 2, st
    assert bits_equal(out_big[0], oracle_merge(big, 60, "sum")[0][0])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_majority_sign_exactly_zero(dtype):
    """n_pos == n_neg: the reference's majority sign is 0 (ties_merging.py:111-118), so `max` multiplies the elements whose
    survivors cancel by 0 (+0) while `sum` / `mean` keep the negative side.  Vector path (whole chunks) and scalar tail."""
    g = torch.Generator().manual_seed(11)
    n = 3 * 8192 + 77
    a = (torch.randint(1, 6, (n,), generator=g).float() * 0.25)
    third = n // 3
    s0, s1 = a.clone(), a.clone()
    s0[third:2 * third] *= -1           # both negative
    s1[third:2 * third] *= -1
    s1[2 * third:] *= -1                # survivors cancel exactly
    s0[2 * third + 100:2 * third + 200] = 0   # ... and some elements with no survivors at all
    s1[2 * third + 100:2 * third + 200] = 0
    srcs = [[s0.to(dtype)], [s1.to(dtype)]]
    dev = [[t.cuda() for t in lst] for lst in srcs]
    for func in ("sum", "mean", "max"):
        out = [torch.full((n,), 7.0, dtype=torch.float32 if func == "mean" else dtype, device="cuda")]
        plan = M.TiesPlan(dev, out)
        plan.run(99, func)
        st = plan.stats()
        want, ost = oracle_merge(srcs, 99, func)
        assert st["majority"] == 0 == int(ost["majority"]) and st["n_ambiguous"] > 0 and st["n_zero"] == 100, st
        assert bits_equal(out[0], want[0]), (func, dtype, int((out[0].cpu().float() != want[0].float()).sum()))
        plan.close()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_sampled_select_bit_exact(dtype):
    """>= 1024 chunks: thresholds come from the sampled bracket + one counting pass — same result as the oracle"""
    g = torch.Generator().manual_seed(21)
    sizes = [6_000_000, 4096 * 11, 3_300_001]
    srcs = [[(torch.randn(n, generator=g) * (0.02 + 0.01 * s)).to(dtype) for n in sizes] for s in range(3)]
    dev = [[t.cuda() for t in lst] for lst in srcs]
    for func, K in (("sum", 20), ("mean", 3), ("max", 65)):
        outs = [torch.empty(n, dtype=torch.float32 if func == "mean" else dtype, device="cuda") for n in sizes]
        plan = M.TiesPlan(dev, outs)
        plan.run(K, func)
        st = plan.stats()
        assert not st["full_select_ran"], st
        want, ost = oracle_merge(srcs, K, func)
        assert st["thresholds"] == [float(x) for x in ost["thresholds"]]
        for o, w in zip(outs, want):
            assert bits_equal(o, w)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n_values,K", [(8, 45), (8, 50), (8, 75), (256, 20), (256, 50.2), (700, 33), (3000, 20)])
def test_counting_pass_every_bracket_width(dtype, n_values, K):
    """The counting pass keeps one SIMD counter per bin boundary for brackets of 1-4 bins and falls back to shared-memory
    bins for wider ones.  Magnitudes drawn uniformly from `n_values` ADJACENT values of the dtype put a chosen share of
    the data in every bin: 8 values -> the bracket is one bin (K = 45) or the two bins around a boundary (K = 50, 75);
    256 / 700 values -> 2-4 bins; 3000 values -> wider than four.  Thresholds and outputs must equal the oracle's."""
    g = torch.Generator().manual_seed(1000 + n_values)
    n = 8192 * 1030 + 77                                   # >= 1024 chunks: the sampled select runs
    base = torch.tensor([1.0], dtype=dtype).view(torch.int16).item()
    srcs = []
    for s in range(2):
        keys = (base + torch.randint(0, n_values, (n,), generator=g)).to(torch.int16)
        sign = torch.randint(0, 2, (n,), generator=g).to(torch.int16) << 15
        srcs.append([(keys | sign).view(dtype)])
    outs = [torch.empty(n, dtype=dtype, device="cuda")]
    plan = M.TiesPlan([[t.cuda() for t in lst] for lst in srcs], outs)
    plan.run(K, "sum")
    st = plan.stats()
    assert not st["full_select_ran"], st
    want, ost = oracle_merge(srcs, K, "sum")
    assert st["thresholds"] == [float(x) for x in ost["thresholds"]], (st["thresholds"], ost["thresholds"])
    assert (st["n_pos"], st["n_neg"], st["n_ambiguous"]) == (ost["n_pos"], ost["n_neg"], ost["ambiguous"])
    assert bits_equal(outs[0], want[0])
    plan.close()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_mean_quotient_every_value_on_device(dtype):
    """`mean` divides by the number of kept entries without a division instruction (q1 = fma(fma(-q0, c, x), r, q0) with the
    hardware's approximate reciprocal r): every finite positive value of the dtype as the kept sum, every count 1..8.
    Source 0 carries the values, c - 1 further sources the smallest subnormal with the same sign (kept: it is their
    threshold; too small to move the 16-bit sum of all but the smallest values), the rest zeros."""
    bits = torch.arange(1, 1 << 15, dtype=torch.int32).to(torch.int16)
    vals = bits.view(dtype)
    vals = vals[torch.isfinite(vals.float())]
    vals = torch.cat([vals, -vals])                        # both elected signs
    tiny = torch.tensor([1], dtype=torch.int16).view(dtype)
    n = vals.numel()
    for c in range(1, 9):
        srcs = [[vals.clone()]]
        for s in range(1, 8):
            t = (tiny.expand(n).clone() * torch.sign(vals.float()).to(dtype)) if s < c else torch.zeros(n, dtype=dtype)
            srcs.append([t])
        outs = [torch.empty(n, dtype=torch.float32, device="cuda")]
        plan = M.TiesPlan([[t.cuda() for t in lst] for lst in srcs], outs)
        plan.run(99.99, "mean")                             # trims the 6 smallest magnitudes only
        want, _ = oracle_merge(srcs, 99.99, "mean")
        assert bits_equal(outs[0], want[0]), (dtype, c, int((outs[0].cpu() != want[0]).sum()))
        assert len(torch.unique(want[0])) > (n * 2) // (3 * c)  # the quotients really cover the value range
        plan.close()


def test_bracket_miss_falls_back_to_full_histogram():
    """data laid out against the sampler (one 512-byte granule per 16 KB chunk is sampled, and exactly those hold zeros):
    the sample says thr = 0, the rank lies outside the bracket, the full-range passes take over — still exact"""
    n_chunks, chunk, gran = 1100, 8192, 256
    g = torch.Generator().manual_seed(22)
    srcs = []
    for s in range(2):
        t = (torch.randn(n_chunks * chunk, generator=g) * 0.02).to(torch.bfloat16)
        v = t.view(n_chunks, chunk // gran, gran)
        for c in range(n_chunks):
            v[c, (c * 7) % 32] = 0
        srcs.append([t])
    outs = [torch.empty(n_chunks * chunk, dtype=torch.bfloat16, device="cuda")]
    plan = M.TiesPlan([[t.cuda() for t in lst] for lst in srcs], outs)
    plan.run(20, "sum")
    st = plan.stats()
    assert st["full_select_ran"], st
    want, ost = oracle_merge(srcs, 20, "sum")
    assert st["thresholds"] == [float(x) for x in ost["thresholds"]]
    assert bits_equal(outs[0], want[0])


def test_unaligned_and_fused_tensors():
    """views at odd offsets take the element path; back-to-back views fuse into one segment — same result either way"""
    g = torch.Generator().manual_seed(8)
    n_src, sizes = 3, [5000, 12288, 3, 8192 * 3]
    slabs = [(torch.randn(sum(sizes) + 1, generator=g) * 0.02).to(torch.bfloat16) for _ in range(n_src)]
    for shift in (0, 1):
        srcs, pos = [[] for _ in range(n_src)], shift
        for n in sizes:
            for s in range(n_src):
                srcs[s].append(slabs[s][pos:pos + n])
            pos += n
        dslabs = [sl.cuda() for sl in slabs]
        dev, pos = [[] for _ in range(n_src)], shift
        for n in sizes:
            for s in range(n_src):
                dev[s].append(dslabs[s][pos:pos + n])
            pos += n
        outs = [torch.empty(n, dtype=torch.float32, device="cuda") for n in sizes]
        plan = M.TiesPlan(dev, outs)
        plan.run(20, "mean")
        plan.stats()
        want, _ = oracle_merge([[t.clone() for t in lst] for lst in srcs], 20, "mean")
        for o, w in zip(outs, want):
            assert bits_equal(o, w)


def test_host_tensor_entry_point_with_caller_outputs():
    """merge.ties_merge_host_tensors (mc_ties_host): new CPU outputs and caller-provided pinned outputs hold the same bits as the
    device-resident plan; wrong outputs are refused."""
    srcs = make_sources(3, [8192, 40000, 17], torch.bfloat16, seed=3, kind="gauss")
    for func in ("sum", "mean"):
        odt = torch.float32 if func == "mean" else torch.bfloat16
        want, ost = oracle_merge(srcs, 20, func)
        got, st = M.ties_merge_host_tensors(srcs, 20, func)
        mine = [torch.empty(t.shape, dtype=odt, pin_memory=True) for t in srcs[0]]
        got2, st2 = M.ties_merge_host_tensors(srcs, 20, func, outputs=mine)
        assert all(a is b for a, b in zip(got2, mine)) and st == st2
        assert st["thresholds"] == [float(x) for x in ost["thresholds"]]
        for t in range(3):
            assert bits_equal(got[t], want[t]) and bits_equal(mine[t], want[t]), (func, t)
    with pytest.raises(ValueError):
        M.ties_merge_host_tensors(srcs, 20, "mean", outputs=[torch.empty(t.shape, dtype=torch.bfloat16) for t in srcs[0]])
    with pytest.raises(ValueError):
        M.ties_merge_host_tensors(srcs, 20, "sum", outputs=[torch.empty(8192, dtype=torch.bfloat16)])


def test_do_merging_matches_reference_fixture(golden):
    g = golden("ties.pt")
    for case in g["vectors"]:
        for f, want in case["outputs"].items():
            got = M.do_merging(case["checks"], K=case["K"], merge_func=f)
            assert list(got) == list(want), case["name"]
            for k in got:
                assert tensor_digest(got[k]) == tensor_digest(want[k]), (case["name"], f, k)


def test_cli_strategies_match_reference_fixture(golden, tmp_path):
    g = golden("ties.pt")
    damc, same = syn.ties_cli_checkpoints()
    dirs = {}
    for fam, ck in (("damc", damc), ("same", same)):
        dirs[fam] = []
        for m in ("vision", "audio"):
            d = str(tmp_path / f"{fam}_{m}")
            syn.save_checkpoint_dir(d, ck[m][0], copy.deepcopy(ck[m][1]))
            dirs[fam].append(d)
    for i, (name, run) in enumerate(g["cli"]["runs"].items()):
        fam, strategy, K = name.split(":")
        out = str(tmp_path / f"out-multimodal-{i}")
        M.main(dirs[fam] + ["-o", out, "--strategy", strategy, "-K", K])
        sd = torch.load(os.path.join(out, "adapter_model.bin"), map_location="cpu")
        assert list(sd) == list(run["digest"]), name
        for k in sd:
            assert tensor_digest(sd[k]) == run["digest"][k], (name, k)
        assert open(os.path.join(out, "config.json")).read() == run["config_json_text"], name
        info = open(os.path.join(out, "merge_info.txt")).read()
        assert info.replace(dirs[fam][0], "{IN0}").replace(dirs[fam][1], "{IN1}").replace(out, "{OUT}") == run["merge_info"], name


METRIC_RTOL = 2e-5  # fp64 reductions here vs the reference's fp32 tree sums


def test_interference_metrics_vs_reference_fixture_and_oracle(golden, tmp_path):
    """calculate_metrics.py: the reference's own numbers on the fixture vectors, the oracle on a sampled-path-sized case,
    and the CLI end to end (merge_info.txt -> merge_metrics.txt)."""
    from modelcompose_b200 import metrics as MT
    g = golden("ties.pt")
    for case in g["vectors"]:
        if "metrics" not in case:
            continue
        keys = sorted(case["checks"][0])
        got = M.interference_metrics_host([[c[k] for k in keys] for c in case["checks"]], 50)
        for k, want in case["metrics"].items():
            assert abs(got[k] - want) <= METRIC_RTOL * max(1.0, abs(want)), (case["name"], k, got[k], want)
    gen = torch.Generator().manual_seed(77)
    srcs = [[(torch.randn(n, generator=gen) * 0.02 + 0.004 * s).to(torch.bfloat16) for n in (5_000_000, 3_700_001)] for s in range(3)]
    plan = M.TiesPlan([[t.cuda() for t in lst] for lst in srcs])   # statistics-only plan: no outputs
    got = plan.metrics(50)
    want = TO.interference_metrics(torch.vstack([torch.cat(lst) for lst in srcs]), 50)
    for k in ("L2", "Cosine", "SSD", "TSSD"):
        assert abs(got[k] - want[k]) <= 5e-6 * max(1.0, abs(want[k])), (k, got[k], want[k])   # same fp32 element ops; fp32 partials over 16 elements, then fp64
    with pytest.raises(Exception, match="without outputs"):
        plan.run(20, "sum")
    with pytest.raises(IndexError):
        M.interference_metrics_host([[torch.zeros(4)]], 50)
    # CLI: ties-mean merge of the two DAMC fixture checkpoints, then calculate_metrics on the output directory
    damc, _ = syn.ties_cli_checkpoints()
    dirs = []
    for m in ("vision", "audio"):
        d = str(tmp_path / f"damc_{m}")
        syn.save_checkpoint_dir(d, damc[m][0], copy.deepcopy(damc[m][1]))
        dirs.append(d)
    out = str(tmp_path / "out-multimodal")
    M.main(dirs + ["-o", out, "--strategy", "ties-mean", "-K", "20"])
    MT.main([out])
    ref_txt = g["cli"]["runs"]["damc:ties-mean:20"]["merge_metrics_txt"]
    got_txt = open(os.path.join(out, "merge_metrics.txt")).read()
    num = lambda line: float(line.split(": ")[1].replace("tensor(", "").replace(")", ""))
    for a, b in zip(got_txt.strip().splitlines(), ref_txt.strip().splitlines()):
        assert a.split(":")[0] == b.split(":")[0] and ("tensor(" in a) == ("tensor(" in b)
        assert abs(num(a) - num(b)) <= 1e-4 * max(1.0, abs(num(b))), (a, b)     # printed with 4 decimals


def test_error_behaviour():
    x = [[torch.zeros(10, dtype=torch.bfloat16, device="cuda")] for _ in range(2)]
    with pytest.raises(Exception, match="float32"):
        M.TiesPlan(x, [torch.zeros(10, dtype=torch.bfloat16, device="cuda")]).run(20, "mean")
    with pytest.raises(RuntimeError, match="kthvalue"):
        M.TiesPlan(x, [torch.zeros(10, dtype=torch.bfloat16, device="cuda")]).run(100, "sum")
    with pytest.raises(ValueError, match="CUDA tensor"):
        M.TiesPlan([[torch.zeros(10, dtype=torch.bfloat16)]], [torch.zeros(10, dtype=torch.bfloat16, device="cuda")])
    with pytest.raises(ValueError, match="Differing parameter names"):
        M.do_merging([{"a": torch.zeros(3)}, {"b": torch.zeros(3)}])


@pytest.mark.parametrize("func", ["sum", "mean", "max"])
def test_full_size_properties(func):
    """3 sources x 160 M bf16 elements (the shared `default` adapters of three vicuna-7B DAMC checkpoints): exact rank of
    the thresholds, census consistency, and outputs re-derived with torch ops on the device from the kernel's own thresholds."""
    n_src, sizes = 3, [4096 * 128, 11008 * 128, 4096 * 128 * 150, 11008 * 128 * 57]   # ~ 160 M elements
    d = sum(sizes)
    g = torch.Generator(device="cuda").manual_seed(123)
    dev = [[(torch.randn(n, generator=g, device="cuda") * 0.02).to(torch.bfloat16) for n in sizes] for _ in range(n_src)]
    odt = torch.float32 if func == "mean" else torch.bfloat16
    outs = [torch.empty(n, dtype=odt, device="cuda") for n in sizes]
    plan = M.TiesPlan(dev, outs)
    plan.run(20, func)
    st = plan.stats()
    k = M.ties_kth_rank(d, 20)
    for s in range(n_src):
        thr = st["thresholds"][s]
        below = sum(int((t.abs().float() < thr).sum()) for t in dev[s])
        at_or_below = sum(int((t.abs().float() <= thr).sum()) for t in dev[s])
        assert below < k <= at_or_below, (s, thr, below, at_or_below, k)      # thr IS the k-th smallest magnitude
    assert st["n_pos"] + st["n_neg"] + st["n_zero"] + st["n_ambiguous"] == d
    assert st["majority"] == (st["n_pos"] > st["n_neg"]) - (st["n_pos"] < st["n_neg"])
    # element-wise re-derivation on the largest tensor (torch ops on the device, fp32 arithmetic as the reference's CPU ops)
    t = 2
    m = [dev[s][t].float() * (dev[s][t].abs().float() >= st["thresholds"][s]).float() for s in range(n_src)]
    total = ((m[0] + m[1]) + m[2]).to(torch.bfloat16).float()
    sg = torch.sign(total)
    sg = torch.where(sg == 0, torch.full_like(sg, float(st["majority"])), sg)
    sel = [x * torch.where(sg > 0, x > 0, x < 0).float() for x in m]
    if func == "max":
        want = (torch.maximum(torch.maximum(sel[0].abs(), sel[1].abs()), sel[2].abs()).to(torch.bfloat16).float() * sg).to(torch.bfloat16)
    else:
        tot = (((torch.zeros_like(sel[0]) + sel[0]) + sel[1]) + sel[2]).to(torch.bfloat16)  # +0 start, as torch's reduction
        want = tot if func == "sum" else tot.float() / torch.clamp(sum((x != 0).float() for x in sel), min=1)
    assert bits_equal(outs[t], want)
