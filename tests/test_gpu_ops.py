"""Operator-level parity: every CUDA kernel family against the CPU oracle / an fp32 torch
reference on the same seeded inputs.  Everything goes through the C ABI (ctypes)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mmu():
    import mmu_b200
    return mmu_b200


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


CASES = [(256, 256, 128), (1000, 520, 200), (72, 104, 96), (136, 768, 3072)]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("M,N,K", CASES)
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm_store(mmu, dtype, tol, M, N, K, a_mn, b_mn):
    A = rnd(M, K, seed=1).to(dtype)
    B = rnd(N, K, seed=2, scale=1 / math.sqrt(K)).to(dtype)
    bias = rnd(N, seed=3)
    ref = A.float() @ B.float().t() + bias
    Ad = (A.t().contiguous() if a_mn else A).cuda()
    Bd = (B.t().contiguous() if b_mn else B).cuda()
    out = mmu.ops.gemm(Ad, Bd, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), bias=bias.cuda(),
                       out_dtype=torch.float32)
    assert rel(out.cpu(), ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_gemm_epilogues(mmu, dtype, tol):
    from oracle import fusion
    M, N, K = 300, 512, 192
    A, B = rnd(M, K, seed=4).to(dtype), rnd(N, K, seed=5, scale=1 / math.sqrt(K)).to(dtype)
    bias, resid = rnd(N, seed=6), rnd(M, N, seed=7)
    z_ref = A.float() @ B.float().t() + bias
    E = mmu._lib
    z = torch.empty(M, N, device="cuda", dtype=dtype)
    u = torch.empty(M, N, device="cuda", dtype=dtype)
    mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_QUICKGELU, out=z, out2=u, bias=bias.cuda())
    assert rel(z.float().cpu(), z_ref) < tol
    assert rel(u.float().cpu(), fusion.quick_gelu(z_ref)) < tol
    out = torch.empty(M, N, device="cuda")
    if dtype == torch.float32:
        mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_RESIDUAL, out=out, bias=bias.cuda(), aux=resid.cuda())
        assert rel(out.cpu(), z_ref + resid) < tol
    else:  # bf16 path: the residual add lives in add_layernorm_fwd, the GEMM mode is rejected
        with pytest.raises(mmu._lib.MMUError):
            mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_RESIDUAL, out=out, bias=bias.cuda(),
                         aux=resid.cuda())
    zz = rnd(M, N, seed=8, scale=2.0).to(dtype)
    s = torch.sigmoid(1.702 * zz.float())
    g_ref = (z_ref - bias) * (s * (1 + 1.702 * zz.float() * (1 - s)))
    out = mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_DGELU, aux=zz.cuda())
    assert rel(out.float().cpu(), g_ref) < max(tol, 1e-4)
    acc = torch.ones(M, N, device="cuda")
    mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_ATOMIC, out=acc, splits=3, alpha=0.5)
    assert rel(acc.cpu(), 1 + 0.5 * (z_ref - bias)) < tol
    # row remap: rows (b, l < 5) of a (B, 5) block land at b*9 + 2 + l
    Am = rnd(40, K, seed=9).to(dtype)
    dst = torch.zeros(8 * 9, N, device="cuda")
    mmu.ops.gemm(Am.cuda(), B.cuda(), out=dst, seg=(5, 9, 2))
    full = Am.float() @ B.float().t()
    got = dst.cpu().view(8, 9, N)[:, 2:7].reshape(40, N)
    assert rel(got, full) < tol
    assert float(dst.cpu().view(8, 9, N)[:, :2].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(1000, 520, 192), (2304, 768, 320), (640, 3072, 768)])
def test_gemm_epilogues_cta_pair_kernel(mmu, M, N, K):
    """The same epilogue modes on the CTA-PAIR kernel (gemm_bf16_tcgen05_kernel<MODE, OBF, 2>,
    tcgen05 cta_group::2, 256x256 tiles): dispatched for unbatched problems with M >= 512 and
    N > 128 (gemm_tcgen05.cu use_pair) -- every big GEMM of the headline configuration.  Shapes
    cover ragged M / N edges (clipped by the output tensor maps), several tiles per cluster and
    both MN-major operand layouts; the reference is fp32 matmul of the bf16-rounded operands, so
    the tolerance only has to cover the bf16 rounding of the OUTPUT (2^-9) and the tanh.approx
    sigmoid (2^-11)."""
    from oracle import fusion
    E = mmu._lib
    dt = torch.bfloat16
    A, B = rnd(M, K, seed=4).to(dt), rnd(N, K, seed=5, scale=1 / math.sqrt(K)).to(dt)
    bias = rnd(N, seed=6)
    acc_ref = A.float() @ B.float().t()
    z_ref = acc_ref + bias
    # QUICKGELU, training form (z and u) and eval form (u only)
    z = torch.empty(M, N, device="cuda", dtype=dt)
    u = torch.empty(M, N, device="cuda", dtype=dt)
    mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_QUICKGELU, out=z, out2=u, bias=bias.cuda())
    assert rel(z.float().cpu(), z_ref) < 4e-3
    assert rel(u.float().cpu(), fusion.quick_gelu(z_ref)) < 5e-3
    u2 = torch.empty(M, N, device="cuda", dtype=dt)
    mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_QUICKGELU, out2=u2, bias=bias.cuda())
    assert torch.equal(u2, u)
    # DGELU with the MN-major weight operand the dgrad GEMM uses (dz = (dx W) * gelu'(z))
    zz = rnd(M, N, seed=8, scale=2.0).to(dt)
    s = torch.sigmoid(1.702 * zz.float())
    g_ref = acc_ref * (s * (1 + 1.702 * zz.float() * (1 - s)))
    out = mmu.ops.gemm(A.cuda(), B.t().contiguous().cuda(), b_mn_major=True, mode=E.EPI_DGELU, aux=zz.cuda())
    assert rel(out.float().cpu(), g_ref) < 5e-3
    out = mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_DGELU, aux=zz.cuda())
    assert rel(out.float().cpu(), g_ref) < 5e-3
    # ATOMIC: split-K accumulation on top of existing contents (fp32, no output rounding), in the
    # wgrad layout (both operands MN-major)
    for splits in (1, 3):
        acc = torch.ones(M, N, device="cuda")
        mmu.ops.gemm(A.t().contiguous().cuda(), B.t().contiguous().cuda(), a_mn_major=True,
                     b_mn_major=True, mode=E.EPI_ATOMIC, out=acc, splits=splits, alpha=0.5)
        assert rel(acc.cpu(), 1 + 0.5 * acc_ref) < 2e-5
    # plain store, fp32 and bf16 outputs
    o32 = mmu.ops.gemm(A.cuda(), B.cuda(), bias=bias.cuda(), out_dtype=torch.float32)
    assert rel(o32.cpu(), z_ref) < 2e-5
    o16 = mmu.ops.gemm(A.cuda(), B.cuda(), bias=bias.cuda())
    assert rel(o16.float().cpu(), z_ref) < 4e-3


@pytest.mark.parametrize("D", [64, 96, 768, 1024])
def test_layernorm(mmu, D):
    from oracle import fusion
    M = 333
    x, g, b = rnd(M, D, seed=1, scale=2.0) + 0.5, 1 + 0.1 * rnd(D, seed=2), 0.1 * rnd(D, seed=3)
    y, mean, rstd = mmu.ops.layernorm_fwd(x.cuda(), g.cuda(), b.cuda())
    assert rel(y.cpu(), fusion.layer_norm(x, g, b)) < 1e-5
    ybf, _, _ = mmu.ops.layernorm_fwd(x.cuda(), g.cuda(), b.cuda(), out_dtype=torch.bfloat16)
    assert rel(ybf.float().cpu(), fusion.layer_norm(x, g, b)) < 1e-2
    dy = rnd(M, D, seed=4)
    xr = x.clone().double().requires_grad_(True)
    gr, br = g.clone().double().requires_grad_(True), b.clone().double().requires_grad_(True)
    fusion.layer_norm(xr, gr, br).backward(dy.double())
    prev = rnd(M, D, seed=5)
    dx, dg, db, dx_lp, cs = mmu.ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, g.cuda(),
                                                  dx=prev.clone().cuda(), want_lp=True,
                                                  want_colsum=True)
    assert rel(dx.cpu(), prev.double() + xr.grad) < 1e-5
    assert rel(dg.cpu(), gr.grad) < 1e-4 and rel(db.cpu(), br.grad) < 1e-4
    assert rel(cs.cpu(), (prev.double() + xr.grad).sum(0)) < 1e-4
    assert rel(dx_lp.float().cpu(), dx.cpu()) < 1e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("B,L,D,H", [(4, 8, 64, 2), (9, 11, 96, 3), (128, 3, 768, 3), (70, 2, 48, 2),
                                     (50, 5, 128, 2), (128, 4, 256, 2), (200, 2, 768, 3), (90, 60, 768, 3),
                                     (33, 7, 256, 1)])
def test_batch_axis_attention(mmu, dtype, tol, B, L, D, H):
    """Against the oracle's restatement of nn.MultiheadAttention(batch_first=False) on (B, L, D)."""
    hd = D // H
    qkv = rnd(B * L, 3 * D, seed=1).to(dtype)
    out, saved = mmu.ops.attention_fwd(qkv.cuda(), B, L, D, H)
    q = qkv.double().requires_grad_(True)
    t = q.view(B, L, 3, H, hd)
    qq, kk, vv = (t[:, :, i].permute(1, 2, 0, 3) for i in range(3))  # (L, H, B, hd)
    s = (qq / math.sqrt(hd)) @ kk.transpose(-1, -2)
    o = (torch.softmax(s, -1) @ vv).permute(2, 0, 1, 3).reshape(B * L, D)
    assert rel(out.float().cpu(), o) < tol
    if saved.dim() == 1:  # SIMT path saves the log-sum-exp; the tensor-core path the bf16 P
        assert rel(saved.cpu().view(L, H, B), torch.logsumexp(s, -1)) < (1e-5 if dtype == torch.float32 else 1e-2)
    else:
        assert rel(saved.float().cpu()[:, :, :B], torch.softmax(s, -1).reshape(L * H, B, B)) < 2e-2
    do = rnd(B * L, D, seed=2).to(dtype)
    o.backward(do.double())
    dqkv = mmu.ops.attention_bwd(qkv.cuda(), out, do.cuda(), saved, B, L, D, H)
    assert rel(dqkv.float().cpu(), q.grad) < tol * 2
    # position-major rows (l*B + b), the engine's layout: same problems, same kernels, other strides
    # -> bit-identical results on the permuted tensors

    def pm(t):
        return t.view(B, L, -1).transpose(0, 1).reshape(L * B, -1).contiguous()

    out_p, saved_p = mmu.ops.attention_fwd(pm(qkv).cuda(), B, L, D, H, pos_major=True)
    assert torch.equal(out_p.cpu(), pm(out.cpu())) and torch.equal(saved_p.cpu(), saved.cpu())
    dqkv_p = mmu.ops.attention_bwd(pm(qkv).cuda(), out_p, pm(do).cuda(), saved_p, B, L, D, H, pos_major=True)
    assert torch.equal(dqkv_p.cpu(), pm(dqkv.cpu()))


@pytest.mark.parametrize("M,D,N", [(300, 192, 520), (1000, 768, 2304), (640, 128, 512)])
def test_gemm_folded_layernorm_and_residual_epilogue(mmu, M, D, N, measured):
    """The eval path's two GEMM epilogues through the C ABI (reference src/model.py:209-212):
    RESID_LN (x' = x + o W^T + b in fp32, raw bf16 copy, per-slab row sums) feeding a GEMM with the
    LayerNorm FOLDED into its epilogue, against LayerNorm -> Linear computed in fp64."""
    E_ = mmu._lib
    bf = torch.bfloat16
    o = rnd(M, D, seed=1).to(bf)
    w_out, b_out = (rnd(D, D, seed=2) / math.sqrt(D)).to(bf), rnd(D, seed=3)
    x = rnd(M, D, seed=4, scale=2.0) + 0.7                      # residual stream with a row mean
    gamma, beta = 1.0 + 0.2 * rnd(D, seed=5), 0.3 * rnd(D, seed=6)
    w_fc, b_fc = rnd(N, D, seed=7) / math.sqrt(D), rnd(N, seed=8)
    nt = (D + 127) // 128
    x1 = torch.empty(M, D, device="cuda")
    raw = torch.empty(M, D, device="cuda", dtype=bf)
    stats = torch.zeros(M, nt, 2, device="cuda")
    mmu.ops.gemm(o.cuda(), w_out.cuda(), mode=E_.EPI_RESID_LN, out=x1, out2=raw, bias=b_out.cuda(),
                 aux=x.cuda(), stats_out=stats)
    x1_ref = x.double() + o.double() @ w_out.double().t() + b_out.double()
    assert rel(x1.cpu(), x1_ref) < 1e-5
    assert torch.equal(raw.cpu(), x1.cpu().to(bf))               # exact copy of what was written
    s_ref = torch.stack([x1.cpu().double().sum(1), (x1.cpu().double() ** 2).sum(1)], 1)
    assert rel(stats.cpu().double().sum(1), s_ref) < 1e-5
    # in-place residual update (aux aliases out) gives the same bits
    x2 = x.cuda().clone()
    mmu.ops.gemm(o.cuda(), w_out.cuda(), mode=E_.EPI_RESID_LN, out=x2, bias=b_out.cuda(), aux=x2)
    assert torch.equal(x2, x1)
    # folded consumer: LayerNorm(x1) W_fc^T + b_fc, plain and with QuickGELU
    wf, cw, bfold = mmu.ops.ln_fold_weights(w_fc.cuda(), gamma.cuda(), beta.cuda(), b_fc.cuda())
    assert rel(wf.float().cpu(), w_fc * gamma) < 5e-3 and rel(cw.cpu(), wf.float().cpu().sum(1)) < 1e-5
    assert rel(bfold.cpu(), b_fc.double() + w_fc.double() @ beta.double()) < 1e-5
    ln = torch.nn.functional.layer_norm(x1_ref, (D,), gamma.double(), beta.double(), 1e-5)
    z_ref = ln @ w_fc.double().t() + b_fc.double()
    z = mmu.ops.gemm(raw, wf, bias=bfold, ln_fold=(stats, cw, 1e-5))
    u = torch.empty(M, N, device="cuda", dtype=bf)
    mmu.ops.gemm(raw, wf, mode=E_.EPI_QUICKGELU, out2=u, bias=bfold, ln_fold=(stats, cw, 1e-5))
    measured("gemm_fold/bf16/z", rel(z.float().cpu(), z_ref))
    assert rel(z.float().cpu(), z_ref) < 1.5e-2
    assert rel(u.float().cpu(), z_ref * torch.sigmoid(1.702 * z_ref)) < 1.5e-2
    # the unfolded path on the same inputs (LayerNorm kernel -> bf16 -> GEMM) is no closer
    h, _, _ = mmu.ops.layernorm_fwd(x1, gamma.cuda(), beta.cuda(), out_dtype=bf)
    z_two = mmu.ops.gemm(h, w_fc.to(bf).cuda(), bias=b_fc.cuda())
    measured("gemm_fold/bf16/z_unfolded", rel(z_two.float().cpu(), z_ref))
    # ln_pre companion: LayerNorm output + raw copy + sums in partial 0
    y, yraw, st = mmu.ops.layernorm_raw_stats(x.cuda(), gamma.cuda(), beta.cuda(), nt)
    y_ref = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)
    assert rel(y.cpu(), y_ref) < 1e-5 and torch.equal(yraw.cpu(), y.cpu().to(bf))
    assert rel(st[:, 0].cpu(), torch.stack([y_ref.sum(1), (y_ref ** 2).sum(1)], 1)) < 1e-4
    assert float(st[:, 1:].abs().max()) == 0.0 if nt > 1 else True


@pytest.mark.parametrize("B,L,D,H", [(128, 5, 768, 3), (100, 3, 768, 3), (37, 4, 512, 2), (128, 160, 768, 3),
                                     (8, 2, 256, 1), (50, 70, 768, 3), (72, 200, 256, 1)])
def test_fused_batch_axis_attention_eval(mmu, B, L, D, H, measured):
    """The fused eval kernel (head_dim 256, B <= 128; keep_probs=False) against the oracle and
    against the two-kernel path, in both row orders; rows / keys beyond B are padding inside the
    kernel and must not leak into the result."""
    hd = D // H
    qkv = rnd(B * L, 3 * D, seed=3).to(torch.bfloat16)
    t = qkv.double().view(B, L, 3, H, hd)
    qq, kk, vv = (t[:, :, i].permute(1, 2, 0, 3) for i in range(3))  # (L, H, B, hd)
    o = (torch.softmax((qq / math.sqrt(hd)) @ kk.transpose(-1, -2), -1) @ vv).permute(2, 0, 1, 3).reshape(B * L, D)
    two, _ = mmu.ops.attention_fwd(qkv.cuda(), B, L, D, H)
    fused, _ = mmu.ops.attention_fwd(qkv.cuda(), B, L, D, H, keep_probs=False)
    err = rel(fused.float().cpu(), o)
    measured("battn_fused/bf16/out", err)
    assert err < 1.5e-2 and rel(two.float().cpu(), o) < 1.5e-2
    assert rel(fused.float().cpu(), two.float().cpu()) < 1.5e-2

    def pm(x):
        return x.view(B, L, -1).transpose(0, 1).reshape(L * B, -1).contiguous()

    fused_p, _ = mmu.ops.attention_fwd(pm(qkv).cuda(), B, L, D, H, pos_major=True, keep_probs=False)
    assert torch.equal(fused_p.cpu(), pm(fused.cpu()))  # same problems, other strides: bit-identical


@pytest.mark.parametrize("N,E,C", [(500, 5, 101), (37, 2, 2), (64, 4, 10), (1001, 1, 101), (33, 3, 300)])
def test_uncertainty_epilogue(mmu, N, E, C):
    from oracle import fusion, uncertainty
    logits = rnd(N, E, C, seed=N, scale=3.0)
    g = torch.Generator().manual_seed(N + 1)
    y = torch.randint(0, C, (N,), generator=g)
    yt = torch.stack([torch.roll(y, e) for e in range(E)], 1).contiguous()  # distinct per head
    # ---- train mode: per-head CE, gradient, per-row accuracy
    dl, pred, scores, accum = mmu.ops.heads_uncertainty_epilogue(
        logits.cuda(), yt.cuda(), 0, grad_scale=1.0 / (N * E), want_grad=True, want_pred=True,
        want_scores=True)
    a = mmu.ops.accum_to_dict(accum)
    z = logits.double().requires_grad_(True)
    loss = fusion.compute_loss(z, yt, eval=False)
    loss.backward()
    assert abs(a["loss_sum"] / a["n_rows"] - float(loss)) < 1e-5 * max(1.0, float(loss))
    assert rel(dl.cpu(), z.grad) < 1e-4
    assert a["n_rows"] == N * E and a["n_samples"] == N
    ref_pred_rows = fusion.predictions(logits, eval=False)
    assert a["n_correct_rows"] == int((ref_pred_rows == yt.reshape(-1)).sum())  # bit-exact
    # ---- scores / histograms against the fp64 definitions
    s = uncertainty.ensemble_scores(logits)
    sc = scores.cpu().double()
    assert torch.equal(pred[:, 1].cpu().long(), s["pred_prob"])  # bit-exact argmax
    assert rel(sc[:, 0], s["conf"]) < 1e-5
    assert float((sc[:, 1] - s["h_pred"]).abs().max()) < 1e-4
    assert float((sc[:, 2] - s["h_exp"]).abs().max()) < 1e-4
    assert float((sc[:, 3] - s["mi"]).abs().max()) < 1e-4
    h = uncertainty.calibration_histograms(logits, y)
    # the kernel's histogram is exactly the binning of the kernel's own scores ...
    cb = uncertainty.bin_index(scores[:, 0].cpu(), 15)
    assert torch.equal(torch.bincount(cb, minlength=15), torch.from_numpy(a["conf_count"]))
    # ... and equals the oracle's unless a score sits within rounding distance of a bin edge
    edge = ((s["conf"] * 15) - (s["conf"] * 15).round()).abs().min()
    if float(edge) > 1e-4:
        assert torch.equal(h["conf_count"], torch.from_numpy(a["conf_count"]))
    # ---- eval mode: CE on the head-mean logits, argmax of the same
    _, pred_e, _, acc_e = mmu.ops.heads_uncertainty_epilogue(logits.cuda(), y.cuda(), 1, want_pred=True)
    ae = mmu.ops.accum_to_dict(acc_e)
    assert abs(ae["loss_sum"] / N - float(fusion.compute_loss(logits.double(), y, eval=True))) < 1e-5
    assert torch.equal(pred_e[:, 0].cpu().long(), s["pred_logit"])
    assert ae["n_correct_rows"] == int((s["pred_logit"] == y).sum())
    assert ae["n_correct_prob"] == int((s["pred_prob"] == y).sum())
    assert torch.equal(torch.from_numpy(ae["conf_correct"]), h["conf_correct"]) or float(edge) <= 1e-4
    ece = uncertainty.ece_from_bins(h["conf_count"], h["conf_correct"], h["conf_sum"])
    m = mmu.metrics.UncertaintyMeter("cuda", C, E)
    m.update(logits.cuda(), y.cuda())
    assert abs(m.compute()["ece"] - ece) < 1e-4


def test_adamw_flat(mmu, golden):
    from oracle import optim
    c = golden("adamw_cosine.pt")
    n = 260  # 257 rounded up to a multiple of 4
    p = torch.zeros(n); p[:257] = c["p0"]
    pd, m, v = p.cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    shadow = torch.zeros(n, dtype=torch.bfloat16).cuda()
    for t, g in enumerate(c["grads"]):
        gg = torch.zeros(n); gg[:257] = g
        mmu.ops.adamw_flat_step(pd, (gg * 4).cuda(), m, v, t + 1, c["lrs"][t], grad_scale=0.25,
                                p_bf16=shadow)
        assert float((pd.cpu()[:257] - c["traj"][t]).abs().max()) < 5e-6  # reference trajectory
    assert float((m.cpu()[:257] - c["exp_avg"]).abs().max()) < 1e-6
    assert torch.equal(shadow.cpu(), pd.cpu().to(torch.bfloat16))


def test_mask_gather(mmu):
    src = rnd(6, 9, 16, seed=3)
    idx = torch.tensor([0, 3, 4, 8], dtype=torch.int32)
    keep = torch.tensor([[1, 1], [0, 1], [1, 0], [1, 1], [0, 0], [1, 1]], dtype=torch.int32)
    out = mmu.ops.mask_gather_tokens(src.cuda(), idx.cuda(), keep.cuda(), modality=0)
    ref = src[:, idx.long()] * keep[:, 0].view(-1, 1, 1)
    assert torch.equal(out.cpu(), ref)  # bit-exact masks / gather
    out = mmu.ops.mask_gather_tokens(src.cuda(), None, keep.cuda(), modality=1, dtype=torch.bfloat16)
    assert torch.equal(out.cpu(), (src * keep[:, 1].view(-1, 1, 1)).to(torch.bfloat16))
    # position-major destination rows (token position, sample): what the engine stages
    out = mmu.ops.mask_gather_tokens(src.cuda(), idx.cuda(), keep.cuda(), modality=0, pos_major=True)
    assert torch.equal(out.cpu(), ref.transpose(0, 1))


def test_posthoc_scoring_matches_notebook_golden(mmu, golden):
    """Device-side post-hoc scoring against the reference notebooks' own functions
    (notebooks/utils.py softmax / get_correlation, food101_robustness.py
    process_predictions_food101) frozen in tests/golden/notebook_scoring.pt."""
    c = golden("notebook_scoring.pt")
    preds, labels = c["preds"], c["labels"].long()          # (S, V, K, C)
    S, V, K, Cn = preds.shape
    n_rep = (V - 3) // 2
    meter = mmu.metrics.PosthocMeter("cuda", n_rep)
    pts = []
    for lo in range(0, S, 16):                               # several batches: sums accumulate
        lg = preds[lo:lo + 16].transpose(0, 1).contiguous().cuda()   # (V, B, K, C)
        pts.append(meter.update(lg, labels[lo:lo + 16].cuda(), want_p_true=True).cpu())
    pt = torch.cat(pts)
    assert rel(pt[:, 0], c["ori"]) < 1e-5 and rel(pt[:, 1], c["image"]) < 1e-5
    assert rel(pt[:, 2], c["text"]) < 1e-5
    assert rel(pt[:, 3:3 + n_rep], c["image_corr"]) < 1e-5
    assert rel(pt[:, 3 + n_rep:], c["text_corr"]) < 1e-5
    out = meter.compute()
    assert out["n_samples"] == S
    assert abs(out["corr_image"] - c["corr_image"]) < 1e-3 * max(abs(c["corr_image"]), 1e-2)
    assert abs(out["corr_text"] - c["corr_text"]) < 1e-3 * max(abs(c["corr_text"]), 1e-2)
    assert out["acc_full"] == pytest.approx(c["acc_full"], abs=1e-9)   # integer counts: exact
    from oracle import uncertainty
    tab = uncertainty.acc_table(preds.numpy(), labels.numpy(), n_rep)
    assert out["acc_image"] == pytest.approx(tab["image"], abs=1e-9)
    assert out["acc_text"] == pytest.approx(tab["text"], abs=1e-9)
    assert out["acc_image_control"] == pytest.approx(float(tab["image_control"].mean()), abs=1e-12)
    assert out["acc_text_control"] == pytest.approx(float(tab["text_control"].mean()), abs=1e-12)


def test_ragged_collator_matches_pad_sequence(mmu, tmp_path):
    """Device-side ragged batch assembly (mmu_ragged_pad via RaggedCollator) against the
    reference's host collate (pad_sequence, src/dataset.py:216-226): bit-exact, zero tails."""
    src = mmu.dataset.SyntheticFlavaDataset(11, l_img=6, l_txt=9, dim=16, num_classes=5, seed=8, ragged=True)
    items = [src[i] for i in range(len(src))]
    mmu.dataset.pack_flava_encodings(((a, b, int(c)) for a, b, c in items), str(tmp_path / "store"))
    ds = mmu.dataset.PackedFlavaDataset(str(tmp_path / "store"))
    collate = mmu.dataset.RaggedCollator(ds, "cuda")
    for idxs in ([0, 1, 2, 3], [10, 4, 7], [5]):
        (gi, gt), gy = collate(idxs)
        (ri, rt), ry = mmu.dataset.collate_fn_flava([items[i] for i in idxs])
        assert torch.equal(gi.cpu(), ri) and torch.equal(gt.cpu(), rt) and torch.equal(gy.cpu(), ry)


@pytest.mark.parametrize("n", [1, 2, 3, 1023, 1024, 1025, 2047, 2049, 4097, 6000])
def test_pair_concordance_counts_bit_exact(mmu, n):
    """mmu_pair_concordance against the oracle's integer pair counts: sizes straddle the 1024-row
    tile, the 2048-column chunk and the diagonal blocks; values are tie-heavy."""
    import numpy as np
    from oracle import rank_stats
    rng = np.random.RandomState(n)
    x = rng.randint(0, 6, size=n).astype(np.float32)
    ys = np.stack([rng.randint(0, 40, size=n).astype(np.float32),
                   rng.randn(n).astype(np.float32), x.copy()])
    got = mmu.ops.pair_concordance(torch.from_numpy(x).cuda(), torch.from_numpy(ys).cuda()).cpu()
    for b in range(3):                                   # x shared by the batch (stride 0)
        assert tuple(int(v) for v in got[b]) == rank_stats.pair_counts(x, ys[b])
    got2 = mmu.ops.pair_concordance(torch.from_numpy(ys).cuda(), torch.from_numpy(ys[[2, 0, 1]].copy()).cuda())
    for b in range(3):
        assert tuple(int(v) for v in got2[b].cpu()) == rank_stats.pair_counts(ys[b], ys[[2, 0, 1][b]])


def test_rank_statistics_match_reference_notebooks(mmu, golden):
    """Device top-5 truncation, Kendall tau-b between heads and AUROC per variant against the
    reference's own notebook functions frozen in tests/golden/rank_stats.pt
    (notebooks/analysis_round_1.py:74-113, notebooks/hatefulmeme_robustness.py:22-41,105-112)."""
    import numpy as np
    c = golden("rank_stats.pt")
    preds, labels = c["predictions"].cuda(), c["labels"].long().cuda()
    for k in range(preds.shape[1]):
        got = mmu.ops.top_truncate(preds[:, k].contiguous(), labels, c["top"], True)
        assert torch.equal(got.cpu().double(), c["muted"][k])                      # bit-exact
    got = mmu.ops.top_truncate(preds[:, 0].contiguous(), None, 3, False)
    assert torch.equal(got.cpu().double(), c["plain_top3"])
    taus = mmu.metrics.head_diversity_kendalltau(preds, labels, top=c["top"])
    assert np.allclose(taus, c["taus"].numpy(), rtol=0, atol=1e-12)
    # AUROC of the notebook's own scores: exact counts -> equal to sklearn's up to its last ulp
    tab = mmu.metrics.auc_table(c["hm_labels"], c["hm_scores"])
    assert np.allclose(tab["AUC"], c["hm_auc"].numpy(), rtol=0, atol=1e-12)
    n = (len(tab["AUC"]) - 3) // 2
    assert tab["image_control"] == pytest.approx(float(c["hm_auc"][3:3 + n].mean()), abs=1e-12)
    # scores computed on device from the (S, V, K, 2) logits: fp32 softmax (max-subtracted) vs the
    # notebook's naive fp32 softmax -> 1e-5 on the scores, AUROC within the reach of re-broken ties
    lg = c["hm_preds"].transpose(0, 1).contiguous().cuda()                         # (V, S, K, 2)
    meter = mmu.metrics.PosthocMeter("cuda", n)
    p1 = meter.class_prob(lg, 1)
    assert rel(p1.cpu(), c["hm_scores"]) < 1e-5
    tab2 = mmu.metrics.auc_table(c["hm_labels"], p1)
    assert np.abs(tab2["AUC"] - c["hm_auc"].numpy()).max() < 2e-3
    assert mmu.metrics.auroc(c["hm_labels"], p1[:, 0].contiguous()) == pytest.approx(tab2["AUC"][0], abs=0)
    # the same table accumulated batch by batch by the sweep's meter (PosthocMeter(auc=True))
    acc_meter = mmu.metrics.PosthocMeter("cuda", n, auc=True)
    for lo in range(0, lg.shape[1], 96):
        acc_meter.update(lg[:, lo:lo + 96].contiguous(), c["hm_labels"][lo:lo + 96].long().cuda())
    res = acc_meter.compute()
    assert np.array_equal(res["auc_per_variant"], tab2["AUC"]) and res["auc_full"] == tab2["full"]
    with pytest.raises(ValueError):
        mmu.metrics.auroc(torch.zeros(8), torch.rand(8))


def test_rank_statistics_full_size_properties(mmu):
    """At evaluation-set sizes (200k scores: O(n^2) = 2e10 pairs on device) the counts obey their
    identities and the AUROC equals sklearn's rank-based value."""
    import numpy as np
    from sklearn.metrics import roc_auc_score
    n = 200_000
    g = torch.Generator().manual_seed(5)
    lab = torch.randint(0, 2, (n,), generator=g)
    sc = (torch.randn(n, generator=g) + 0.5 * lab).mul(64).round().div(64)          # ties
    a = mmu.metrics.auroc(lab, sc)
    assert abs(a - roc_auc_score(lab.numpy(), sc.numpy())) < 1e-12
    assert abs(mmu.metrics.auroc(lab, -sc) - (1.0 - a)) < 1e-15
    cnt = mmu.metrics.pair_counts(lab, sc)[0]
    n_pos = int(lab.sum())
    tot = n * (n - 1) // 2
    joint = int(cnt.sum()) - tot
    assert int(cnt[2]) == n_pos * (n_pos - 1) // 2 + (n - n_pos) * (n - n_pos - 1) // 2
    assert int(cnt[0]) + int(cnt[1]) + int(cnt[3]) - joint == n_pos * (n - n_pos)
    assert mmu.metrics.kendalltau(sc, sc) == pytest.approx(1.0, abs=1e-15)
    assert mmu.metrics.kendalltau(sc, -sc) == pytest.approx(-1.0, abs=1e-15)


@pytest.mark.parametrize("N,E,C", [(777, 5, 101), (130, 4, 10), (64, 6, 101), (99, 3, 2), (50, 2, 60)])
def test_uncertainty_epilogue_batched_heads_bit_identical(mmu, N, E, C, monkeypatch):
    """Eval mode reduces HB heads together (ce_uncertainty_kernel<..., HB>); every HB that divides E
    must give the head-by-head kernel's results bit for bit (scores, predictions, every bin)."""
    logits = rnd(N, E, C, seed=N + E, scale=4.0).cuda()
    g = torch.Generator().manual_seed(N)
    y = torch.randint(0, C, (N,), generator=g).cuda()
    outs = {}
    for hb in [h for h in (1, 2, 3, 4, 5) if E % h == 0]:
        monkeypatch.setenv("MMU_CE_HB", str(hb))
        _, pred, scores, accum = mmu.ops.heads_uncertainty_epilogue(logits, y, 1, want_pred=True,
                                                                    want_scores=True)
        outs[hb] = (pred.cpu(), scores.cpu(), accum.cpu())
    monkeypatch.delenv("MMU_CE_HB")
    _, pred, scores, accum = mmu.ops.heads_uncertainty_epilogue(logits, y, 1, want_pred=True,
                                                                want_scores=True)
    outs["default"] = (pred.cpu(), scores.cpu(), accum.cpu())
    assert len(outs) >= 3
    for hb, (p, s, a) in outs.items():
        assert torch.equal(p, outs[1][0]) and torch.equal(s, outs[1][1]), hb
        # integer words of the accumulator are order independent; the fp64 sums are atomics
        ia = a.view(torch.int64)[:mmu._lib.ACC_INT_WORDS]
        assert torch.equal(ia, outs[1][2].view(torch.int64)[:mmu._lib.ACC_INT_WORDS]), hb


@pytest.mark.parametrize("mode", ["random", "guided"])
@pytest.mark.parametrize("B,p", [(128, 0.5), (77, 0.25), (300, 1.0), (16, 0.0)])
def test_modality_keep_mask_device_equals_oracle(mmu, mode, B, p):
    """mmu_modality_keep_mask (device, scores never leave the GPU) against the oracle's
    sample-by-sample definition (oracle/shaping.py) under the same host generator: bit-exact."""
    from oracle import shaping
    g = torch.Generator().manual_seed(B)
    scores = torch.rand(B, 4, generator=g)
    scores[::7, 1] = scores[::7, 0]                       # exact ties -> image is dropped
    ref = shaping.modality_dropout_mask(B, p, mode, scores[:, :2], torch.Generator().manual_seed(11))
    sd = scores.cuda()
    got = mmu.robustness.modality_dropout_mask_device(B, p, mode, "cuda", sd[:, 0], sd[:, 1],
                                                      torch.Generator().manual_seed(11))
    assert got.dtype == torch.int32 and torch.equal(got.cpu(), ref)
    host = mmu.robustness.modality_dropout_mask(B, p, mode, scores[:, :2], torch.Generator().manual_seed(11))
    assert torch.equal(host, ref)
    if p == 1.0:
        assert int((ref.sum(1) == 1).sum()) == B           # every sample lost exactly one modality
