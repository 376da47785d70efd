"""Parity of the CUDA path against the CPU oracle AT THE BENCHMARKED SIZE (BASELINE.json
configs[1] / configs[2] as bench.py runs them): B = 128 samples, 197 image + 40 text tokens of
768 features, D = 768, 3 layers, 3 heads (hd = 256), E = 5 heads, C = 101 classes -> M = 30 336
rows per GEMM, every launch on the CTA-pair (cta_group::2) tcgen05 kernel with many tiles per
CTA.  Reference arithmetic: src/model.py:258-304 (forward, loss), train.py:196-202 (AdamW).

The oracle (oracle/fusion.py, explicit fp32 torch-CPU tensor arithmetic pinned to goldens of the
unmodified reference in tests/test_oracle_golden.py) runs the SAME seeded inputs; one
forward+backward costs ~10-20 s of host time at this size, so one oracle pass is shared by the
fp32 and bf16 comparisons.

Tolerances
* fp32 engine: <= 1e-3 relative on logits / loss / every gradient (north_star), bit-exact argmax,
  bit-exact histogram bins wherever the oracle's score is not within 1e-4 of a bin edge.
* bf16 engine (bf16 GEMM operands, fp32 accumulation, fp32 residual stream / LayerNorm /
  softmax statistics): the MEASURED errors are printed by every test (`pytest -s`) and recorded
  in DESIGN.md section 2; the bounds below are <= 3x those measurements.  Predictions of a
  bf16 path cannot be bit-exact against an fp32 reference when two logits are closer than the
  bf16 error; the flip COUNT is measured, printed and bounded instead.
"""
import math
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(B=128, l_img=197, l_txt=40, D=768, heads=3, layers=3, E=5, C=101)

# bf16 path at this size: bounds (<= 3x the measured values printed by the tests)
# measured on a B200 (round 2): logits 2.9e-3 .. 3.0e-3, loss 8.4e-5, worst gradient 7.4e-3 (1.28e-2
# with a modality mask), 2 of 640 per-head argmax flips, 7 / 6 of 1 280 sweep predictions (mean
# logits / mean probabilities), 5 of 1 280 confidence bins moved, ECE 0.17927 vs 0.17929
BF16_LOGIT_TOL = 9e-3        # max |dlogit| / max |logit|
BF16_LOSS_TOL = 2.5e-4       # relative
BF16_GRAD_TOL = 3e-2         # max |dg| / max |g| per tensor
BF16_MAX_ROW_FLIPS = 6       # of B*E = 640 per-head argmax rows (train protocol)
BF16_MAX_SWEEP_FLIPS = 20    # of 1 280 (level, sample) head-mean predictions
BF16_MAX_BIN_MOVES = 15      # samples whose confidence bin differs, of 1 280


@pytest.fixture(scope="module")
def mmu():
    import mmu_b200
    return mmu_b200


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def make_model(mmu, precision, E=CFG["E"], avg_pool=False, seed=42):
    torch.manual_seed(seed)
    m = mmu.FlavaFusionTransfomer(out_dim=E, num_classes=CFG["C"],
                                  multimodal_num_attention_heads=CFG["heads"],
                                  multimodal_num_hidden_layers=CFG["layers"], drop=0.0,
                                  avg_pool=avg_pool, precision=precision)
    with torch.no_grad():
        # default init leaves the 101 logits of a row within ~0.1 of each other; scale the heads
        # so that predictions have margins comparable to a trained model's (|logit| ~ 3)
        for e in range(E):
            m.output_layers[e].weight.mul_(8.0)
    return m


@pytest.fixture(scope="module")
def case(mmu):
    """Seeded inputs, parameters and ONE oracle forward+backward at the headline size."""
    from oracle import fusion
    torch.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0))))
    m = make_model(mmu, "fp32")
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(CFG["B"], CFG["l_img"], CFG["D"], generator=g)
    txt = torch.randn(CFG["B"], CFG["l_txt"], CFG["D"], generator=g)
    y = torch.randint(0, CFG["C"], (CFG["B"],), generator=g)
    yt = y.unsqueeze(1).repeat(1, CFG["E"])
    t0 = time.perf_counter()
    logits, loss, grads = fusion.loss_and_grads(P, (img, txt), yt, CFG["heads"], False)
    print(f"\n[headline] oracle forward+backward at B=128, L=237, D=768: {time.perf_counter() - t0:.1f} s")
    return dict(P=P, img=img, txt=txt, y=y, yt=yt, logits=logits, loss=loss, grads=grads)


def run_engine(mmu, case, precision, keep=None):
    m = make_model(mmu, precision)
    m.load_state_dict(case["P"], strict=True)
    m.cuda().train()
    opt = mmu.FusedAdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-3)
    opt.zero_grad()
    x = (case["img"].cuda(), case["txt"].cuda())
    logits = m(x) if keep is None else m(x, keep_mask=keep.cuda())
    loss = m.compute_loss(logits, case["yt"].cuda())
    loss.backward()
    return m, opt, logits.detach().cpu(), float(loss.detach())


def grad_errors(m, ref_grads):
    out = {}
    for k, p in m.named_parameters():
        g = ref_grads[k]
        scale = float(g.abs().max())
        if scale < 1e-9:
            out[k] = (float(p.grad.abs().max()), 0.0)       # (absolute, 0) for dead tensors
        else:
            out[k] = (float((p.grad.cpu() - g).abs().max()) / scale, scale)
    return out


def test_headline_fp32_engine_vs_oracle(mmu, case):
    """fp32 engine: logits / loss / EVERY gradient <= 1e-3 relative, argmax bit-exact, one AdamW step."""
    from oracle import optim
    m, opt, logits, loss = run_engine(mmu, case, "fp32")
    e_logit = rel(logits, case["logits"])
    e_loss = abs(loss - float(case["loss"])) / abs(float(case["loss"]))
    errs = grad_errors(m, case["grads"])
    worst = max((v[0], k) for k, v in errs.items() if v[1] > 0)
    print(f"\n[headline fp32] logits {e_logit:.2e}  loss {e_loss:.2e}  worst grad {worst[0]:.2e} ({worst[1]})")
    assert e_logit < 1e-3 and e_loss < 1e-3
    assert torch.equal(logits.argmax(-1), case["logits"].argmax(-1))           # bit-exact
    acc = float(mmu.acc(logits.cuda(), case["yt"].cuda(), False, True))
    assert acc == pytest.approx(float((case["logits"].argmax(-1) == case["yt"]).float().mean() * 100), abs=1e-4)
    for k, (e, scale) in errs.items():
        if scale == 0.0:
            assert e < 1e-9, k          # text projection: dead without avg_pool, stays exactly dead
        else:
            assert e < 1e-3, (k, e)
    opt.step()
    for k, p in m.named_parameters():
        g = case["grads"][k]
        ref_p, _, _ = optim.adamw_step(case["P"][k], g, torch.zeros_like(g), torch.zeros_like(g), 1, 1e-3)
        ok = g.abs() > 1e-3 * g.abs().max().clamp_min(1e-12)   # Adam's first step is lr*sign(g): only
        d = (p.detach().cpu() - ref_p).abs()                  # well-conditioned elements compare
        assert float(d[ok].max() if ok.any() else 0.0) < 2e-5, k
        assert float(d.max()) < 2.1e-3, k


def test_headline_bf16_engine_vs_oracle(mmu, case):
    """bf16 tensor-core engine against the SAME fp32 oracle pass: measured error printed, bounded
    at <= 3x; per-head argmax flip count printed and bounded."""
    m, _, logits, loss = run_engine(mmu, case, "bf16")
    e_logit = rel(logits, case["logits"])
    e_loss = abs(loss - float(case["loss"])) / abs(float(case["loss"]))
    errs = grad_errors(m, case["grads"])
    worst = max((v[0], k) for k, v in errs.items() if v[1] > 0)
    flips = int((logits.argmax(-1) != case["logits"].argmax(-1)).sum())
    # a flip is only legitimate where the oracle's own top-2 margin is below the bf16 logit error
    top2 = case["logits"].topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    bad = (logits.argmax(-1) != case["logits"].argmax(-1)) & (margin > 2 * e_logit * case["logits"].abs().max())
    print(f"\n[headline bf16] logits {e_logit:.2e}  loss {e_loss:.2e}  worst grad {worst[0]:.2e} "
          f"({worst[1]})  argmax flips {flips}/{logits.shape[0] * logits.shape[1]}")
    assert e_logit < BF16_LOGIT_TOL and e_loss < BF16_LOSS_TOL
    for k, (e, scale) in errs.items():
        if scale == 0.0:
            assert e < 1e-9, k
        else:
            assert e < BF16_GRAD_TOL, (k, e)
    assert flips <= BF16_MAX_ROW_FLIPS
    assert int(bad.sum()) == 0          # every flip sits inside the error band


def sweep_variants(mmu):
    torch.manual_seed(77)
    return [mmu.robustness.mask_level_variant(CFG["l_img"], CFG["l_txt"], "image", k, 10) for k in range(10)]


@pytest.fixture(scope="module")
def sweep_oracle(mmu, case):
    """One oracle forward per mask level (reference loop eval_transformer_robustness.py:99-125)."""
    from oracle import fusion, shaping
    variants = sweep_variants(mmu)
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = torch.stack([fusion.flava_fusion_forward(
            case["P"], shaping.apply_variant(case["img"], case["txt"], v), CFG["heads"], False)
            for v in variants])
    print(f"\n[headline] oracle 10-level sweep: {time.perf_counter() - t0:.1f} s")
    return variants, ref


def test_packed_sweep_is_bit_reproducible_at_headline_size(mmu, case):
    """The bf16 eval forward has no atomics (residual / LayerNorm epilogues, fused batch-axis
    attention, per-slab row sums): twelve runs of the packed 10-level sweep at B = 128, 1 327
    positions -- with other work queued in between to move the timing -- must agree bit for bit.
    A missing barrier or a staging box recycled too early in the new kernels would show up here."""
    m = make_model(mmu, "bf16")
    m.load_state_dict(case["P"], strict=True)
    m.cuda().eval()
    variants = sweep_variants(mmu)
    img, txt = case["img"].cuda(), case["txt"].cuda()
    noise = torch.randn(4096, 4096, device="cuda")
    with torch.no_grad():
        first = m.forward_variants((img, txt), variants).clone()
        for i in range(11):
            if i % 3 == 0:
                noise = noise @ noise * 1e-4       # unrelated work in the queue
            again = m.forward_variants((img, txt), variants)
            assert torch.equal(again, first), f"run {i + 1} differs from run 0"
    assert torch.isfinite(first).all()


def _sweep_compare(mmu, case, sweep_oracle, precision):
    from oracle import uncertainty
    variants, ref = sweep_oracle
    m = make_model(mmu, precision)
    m.load_state_dict(case["P"], strict=True)
    m.cuda().eval()
    with torch.no_grad():
        got = m.forward_variants((case["img"].cuda(), case["txt"].cuda()), variants)
    flat, labels = got.view(-1, CFG["E"], CFG["C"]), case["y"].repeat(10)
    _, pred, scores, accum = mmu.ops.heads_uncertainty_epilogue(flat, labels.cuda(), 1, want_pred=True,
                                                                want_scores=True)
    a = mmu.ops.accum_to_dict(accum)
    ref_flat = ref.view(-1, CFG["E"], CFG["C"])
    h = uncertainty.calibration_histograms(ref_flat, labels)
    s = uncertainty.ensemble_scores(ref_flat)
    e_logit = rel(got.cpu(), ref)
    flips_logit = int((pred[:, 0].cpu().long() != s["pred_logit"]).sum())
    flips_prob = int((pred[:, 1].cpu().long() != s["pred_prob"]).sum())
    bins_engine = uncertainty.bin_index(scores[:, 0].cpu(), 15)
    bins_oracle = uncertainty.bin_index(s["conf"], 15)
    moved = int((bins_engine != bins_oracle).sum())
    near_edge = ((s["conf"] * 15) - (s["conf"] * 15).round()).abs() < 1e-4
    sc = scores.cpu().double()
    e_scores = dict(conf=rel(sc[:, 0], s["conf"]),
                    h_pred=float((sc[:, 1] - s["h_pred"]).abs().max()),
                    h_exp=float((sc[:, 2] - s["h_exp"]).abs().max()),
                    mi=float((sc[:, 3] - s["mi"]).abs().max()))
    ece_o = uncertainty.ece_from_bins(h["conf_count"], h["conf_correct"], h["conf_sum"])
    ece_e = uncertainty.ece_from_bins(torch.from_numpy(a["conf_count"]), torch.from_numpy(a["conf_correct"]),
                                      torch.from_numpy(a["conf_sum"]))
    print(f"\n[headline sweep {precision}] logits {e_logit:.2e}  pred flips logit/prob {flips_logit}/{flips_prob} "
          f"of {flat.shape[0]}  conf bins moved {moved} (near edge {int(near_edge.sum())})  scores {e_scores}  "
          f"ECE {ece_e:.5f} vs {ece_o:.5f}")
    # the kernel's histogram is exactly the binning of the kernel's own scores
    assert torch.equal(torch.bincount(bins_engine, minlength=15), torch.from_numpy(a["conf_count"]))
    assert a["n_samples"] == flat.shape[0]
    return dict(e_logit=e_logit, flips_logit=flips_logit, flips_prob=flips_prob, moved=moved,
                near_edge=int(near_edge.sum()), bins_engine=bins_engine, bins_oracle=bins_oracle,
                scores=e_scores, ece=(ece_e, ece_o), a=a, h=h, moved_mask=(bins_engine != bins_oracle),
                near_mask=near_edge)


def test_headline_packed_sweep_fp32_vs_oracle(mmu, case, sweep_oracle):
    """Packed 10-level sweep (one pass over 1 327 positions x 128 samples) against one oracle
    forward per level: logits <= 1e-3, predictions bit-exact, confidence-bin counts bit-exact
    (samples within 1e-4 of a bin edge excepted), entropy / MI / ECE within 1e-3."""
    r = _sweep_compare(mmu, case, sweep_oracle, "fp32")
    assert r["e_logit"] < 1e-3
    assert r["flips_logit"] == 0 and r["flips_prob"] == 0
    assert not bool((r["moved_mask"] & ~r["near_mask"]).any())
    if r["near_edge"] == 0:
        assert torch.equal(torch.from_numpy(r["a"]["conf_count"]), r["h"]["conf_count"])
        assert torch.equal(torch.from_numpy(r["a"]["conf_correct"]), r["h"]["conf_correct"])
    assert r["scores"]["conf"] < 1e-3 and max(r["scores"]["h_pred"], r["scores"]["h_exp"]) < 1e-3 * math.log(CFG["C"])
    assert r["scores"]["mi"] < 1e-3
    assert abs(r["ece"][0] - r["ece"][1]) < 1e-3


def test_headline_packed_sweep_bf16_vs_oracle(mmu, case, sweep_oracle):
    r = _sweep_compare(mmu, case, sweep_oracle, "bf16")
    assert r["e_logit"] < BF16_LOGIT_TOL
    assert r["flips_logit"] <= BF16_MAX_SWEEP_FLIPS and r["flips_prob"] <= BF16_MAX_SWEEP_FLIPS
    assert r["moved"] <= BF16_MAX_BIN_MOVES
    assert abs(r["ece"][0] - r["ece"][1]) < 1e-3


@pytest.mark.parametrize("mode", ["random", "guided"])
def test_headline_keep_mask_vs_oracle(mmu, case, mode):
    """Modality dropout through the MODEL at the headline size (keep_mask zero-fills a modality
    per sample inside the stem's gather kernel): fp32 engine forward + backward against the oracle
    run on inputs the oracle masks itself, and the bf16 engine within its stated tolerance."""
    from oracle import fusion, shaping
    gen = torch.Generator().manual_seed(5)
    scores = torch.rand(CFG["B"], 2, generator=gen)
    keep = mmu.robustness.modality_dropout_mask(CFG["B"], 0.5, mode, scores, torch.Generator().manual_seed(6))
    keep_o = shaping.modality_dropout_mask(CFG["B"], 0.5, mode, scores, torch.Generator().manual_seed(6))
    assert torch.equal(keep, keep_o) and 0 < int((keep[:, 0] == 0).sum()) < CFG["B"]
    img_o, txt_o = shaping.apply_keep_mask(case["img"], case["txt"], keep_o)
    ref_logits, ref_loss, ref_grads = fusion.loss_and_grads(case["P"], (img_o, txt_o), case["yt"],
                                                            CFG["heads"], False)
    assert rel(ref_logits, case["logits"]) > 1e-2          # the mask changes the result
    for precision, tl, tg in (("fp32", 1e-3, 1e-3), ("bf16", BF16_LOGIT_TOL, BF16_GRAD_TOL)):
        m, _, logits, loss = run_engine(mmu, case, precision, keep=keep)
        e_logit = rel(logits, ref_logits)
        errs = grad_errors(m, ref_grads)
        worst = max((v[0], k) for k, v in errs.items() if v[1] > 0)
        print(f"\n[headline keep_mask {mode} {precision}] logits {e_logit:.2e}  worst grad {worst[0]:.2e} ({worst[1]})")
        assert e_logit < tl and abs(loss - float(ref_loss)) < (1e-3 if precision == "fp32" else BF16_LOSS_TOL) * abs(float(ref_loss))
        if precision == "fp32":
            assert torch.equal(logits.argmax(-1), ref_logits.argmax(-1))
        for k, (e, scale) in errs.items():
            assert (e < 1e-9) if scale == 0.0 else (e < tg), (k, e)


def test_headline_avg_pool_all_tokens_live(mmu):
    """avg_pool=True (E = 2: image head / text head read the MEAN over their tokens,
    src/model.py:282-284): every one of the 30 336 rows is live in the forward AND the backward,
    so this is the case that holds the full-size GEMM / LayerNorm / attention backward to the
    oracle on all rows (the E = 5 model above only back-propagates through token positions < 5)."""
    from oracle import fusion
    m32 = make_model(mmu, "fp32", E=2, avg_pool=True, seed=7)
    P = {k: v.detach().clone() for k, v in m32.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    img = torch.randn(CFG["B"], CFG["l_img"], CFG["D"], generator=g)
    txt = torch.randn(CFG["B"], CFG["l_txt"], CFG["D"], generator=g)
    yt = torch.randint(0, CFG["C"], (CFG["B"],), generator=g).unsqueeze(1).repeat(1, 2)
    ref_logits, ref_loss, ref_grads = fusion.loss_and_grads(P, (img, txt), yt, CFG["heads"], True)
    c = dict(P=P, img=img, txt=txt, yt=yt)
    for precision, tl, tg in (("fp32", 1e-3, 1e-3), ("bf16", BF16_LOGIT_TOL, BF16_GRAD_TOL)):
        m = make_model(mmu, precision, E=2, avg_pool=True, seed=7)
        m.load_state_dict(P, strict=True)
        m.cuda().train()
        m.zero_grad()
        logits = m((img.cuda(), txt.cuda()))
        loss = m.compute_loss(logits, yt.cuda())
        loss.backward()
        e_logit = rel(logits.detach().cpu(), ref_logits)
        errs = grad_errors(m, ref_grads)
        worst = max((v[0], k) for k, v in errs.items() if v[1] > 0)
        print(f"\n[headline avg_pool {precision}] logits {e_logit:.2e}  loss "
              f"{abs(float(loss.detach()) - float(ref_loss)) / float(ref_loss):.2e}  worst grad {worst[0]:.2e} ({worst[1]})")
        assert e_logit < tl
        assert abs(float(loss.detach()) - float(ref_loss)) < (1e-3 if precision == "fp32" else BF16_LOSS_TOL) * float(ref_loss)
        if precision == "fp32":
            assert torch.equal(logits.detach().cpu().argmax(-1), ref_logits.argmax(-1))
        for k, (e, scale) in errs.items():
            assert scale > 0 and e < tg, (k, e)      # nothing is dead with avg_pool
    del c


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_headline_live_token_path_is_bit_identical(mmu, case, precision):
    """``live_tokens=True`` computes only the token positions that can reach a head (positions
    < E without avg_pool, src/model.py:286-287) -- 5 of 237 here.  Same kernels on the same rows:
    train-mode logits, eval logits and the packed 10-level sweep are BIT-identical to the
    as-written path; every gradient agrees to the summation order of the split-K weight-gradient
    atomics (the skipped rows contribute exact zeros), and the dead text projection stays at 0."""
    outs = {}
    variants = sweep_variants(mmu)
    for live in (False, True):
        torch.manual_seed(42)
        m = mmu.FlavaFusionTransfomer(out_dim=CFG["E"], num_classes=CFG["C"],
                                      multimodal_num_attention_heads=CFG["heads"],
                                      multimodal_num_hidden_layers=CFG["layers"], drop=0.0,
                                      avg_pool=False, precision=precision, live_tokens=live)
        m.load_state_dict(case["P"], strict=True)
        m.cuda().train()
        m.zero_grad()
        x = (case["img"].cuda(), case["txt"].cuda())
        logits = m(x)
        m.compute_loss(logits, case["yt"].cuda()).backward()
        grads = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters()}
        m.eval()
        with torch.no_grad():
            ev = m(x).cpu()
            sw = m.forward_variants(x, variants).cpu()
            keep = torch.ones(CFG["B"], 2, dtype=torch.int32)
            keep[::3, 0] = 0
            km = m(x, keep_mask=keep.cuda()).cpu()
        outs[live] = (logits.detach().cpu(), grads, ev, sw, km)
    a, b = outs[False], outs[True]
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    worst = 0.0
    for k in a[1]:
        scale = float(a[1][k].abs().max())
        if scale == 0.0:
            assert float(b[1][k].abs().max()) == 0.0, k
        else:
            worst = max(worst, float((a[1][k] - b[1][k]).abs().max()) / scale)
    print(f"\n[headline live tokens {precision}] logits bit-identical; worst gradient difference {worst:.2e} of max|g|")
    assert worst < 2e-5
