"""Dropout with p > 0 in training (reference: ``nn.Dropout(drop)`` between ``c_fc`` and QuickGELU,
src/model.py:195-201; ``ImageBertEmbeddings.dropout`` src/mmbt.py:56,82 and the hidden /
attention-probability dropouts inside pytorch_pretrained_bert's ``BertModel`` for MMBT).

The reference draws its masks from torch's Philox stream; no other generator can reproduce those
draws, so parity is twofold (oracle/dropout.py):
  * STATISTICAL against nn.Dropout's definition: independent keeps with probability 1 - p, scale
    1 / (1 - p), eval mode untouched, p = 0 bit-identical to the dropout-free path;
  * BIT-EXACT masks against the oracle's integer restatement of the engine's counter-based mask
    function, which lets the oracle apply the very same masks: logits / loss / every gradient then
    hold at the usual tolerances (fp32 1e-3; bf16 as stated in test_gpu_model.py).
"""
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


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 5e-3)])
@pytest.mark.parametrize("M,N,K", [(300, 520, 192), (1000, 768, 256)])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_gemm_epilogue_dropout_masks_are_the_oracles(mmu, dtype, tol, M, N, K, p):
    """QUICKGELU / DGELU epilogues with dropout (single-CTA and CTA-pair tcgen05 kernels, fp32
    FFMA kernel): the dropped elements are EXACTLY the oracle's (integer hash restated in numpy),
    the survivors are scaled by 1 / (1 - p), u = QuickGELU(dropout(z)), and the backward epilogue
    regenerates the same mask."""
    from oracle import dropout, fusion
    E = mmu._lib
    seed, site = 0x1234567 + M, 3
    A = rnd(M, K, seed=1).to(dtype)
    B = rnd(N, K, seed=2, scale=1 / math.sqrt(K)).to(dtype)
    bias = rnd(N, seed=3) + 0.3              # keeps |z| away from 0 so that zeros mean "dropped"
    z_ref = A.float() @ B.float().t() + bias
    mult = dropout.multiplier(p, seed, site, M * N, (M, N))
    z = torch.empty(M, N, device="cuda", dtype=dtype)
    u = torch.empty(M, N, device="cuda", dtype=dtype)
    mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_QUICKGELU, out=z, out2=u, bias=bias.cuda(),
                 dropout=(p, seed, site))
    zc = z.float().cpu()
    live = z_ref.abs() > 1e-2
    assert torch.equal((zc != 0)[live], (mult != 0)[live])                     # bit-exact mask
    assert rel(zc, z_ref * mult) < tol
    assert rel(u.float().cpu(), fusion.quick_gelu(z_ref * mult)) < max(tol, 1e-4)
    kept = float((mult != 0).float().mean())
    assert abs(kept - (1 - p)) < 4 * math.sqrt(p * (1 - p) / (M * N))         # binomial 4 sigma
    # backward: dz = acc * gelu'(z_saved) * mask / (1 - p)
    zz = (rnd(M, N, seed=8, scale=2.0)).to(dtype)
    s = torch.sigmoid(1.702 * zz.float())
    g_ref = (z_ref - bias) * (s * (1 + 1.702 * zz.float() * (1 - s))) * mult
    out = mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_DGELU, aux=zz.cuda(), dropout=(p, seed, site))
    assert rel(out.float().cpu(), g_ref) < max(tol, 1e-4)
    assert torch.equal((out.float().cpu() != 0)[live & (g_ref.abs() > 1e-3)], (mult != 0)[live & (g_ref.abs() > 1e-3)])
    # another site / seed gives another mask
    assert not torch.equal(dropout.keep_mask(p, seed, site + 1, M * N), dropout.keep_mask(p, seed, site, M * N))
    with pytest.raises(mmu._lib.MMUError):
        mmu.ops.gemm(A.cuda(), B.cuda(), mode=E.EPI_QUICKGELU, out=z, out2=u, dropout=(1.0, seed, site))


def _build(mmu, cfg, sd, precision, drop):
    klass = mmu.FlavaFusionTransfomerwithCLSToken if cfg["cls"] else mmu.FlavaFusionTransfomer
    m = klass(out_dim=cfg["E"], num_classes=cfg["C"], image_hidden_size=cfg["d_img"],
              text_hidden_size=cfg["d_txt"], multimodal_hidden_size=cfg["D"],
              multimodal_num_attention_heads=cfg["heads"], multimodal_num_hidden_layers=cfg["layers"],
              drop=drop, avg_pool=cfg["avg_pool"], precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.mark.parametrize("precision,tl,tg", [("fp32", 1e-3, 1e-3), ("bf16", 2.5e-2, 4e-2)])
@pytest.mark.parametrize("name", ["plain_E2", "avgpool_E2", "cls_E3"])
def test_flava_training_with_dropout_matches_oracle_under_the_same_masks(mmu, golden, name, precision, tl, tg, measured):
    """drop = 0.25: train-mode forward + backward of the engine against the oracle applying the
    masks the engine drew (seed read back from the model); eval mode ignores dropout; a second
    forward draws a different mask; torch.manual_seed reproduces a run."""
    from oracle import fusion
    c = golden("flava_small.pt")[name]
    cfg = c["cfg"]
    p = 0.25
    m = _build(mmu, cfg, c["state_dict"], precision, p).train()
    x = (c["img"].cuda(), c["txt"].cuda())
    torch.manual_seed(77)
    m.zero_grad()
    logits = m(x)
    seed = m.last_dropout_seed
    loss = m.compute_loss(logits, c["y_train"].cuda())
    loss.backward()
    ref_logits, ref_loss, ref_grads = fusion.loss_and_grads(c["state_dict"], (c["img"], c["txt"]), c["y_train"],
                                                            cfg["heads"], cfg["avg_pool"], dropout=(p, seed))
    assert rel(ref_logits, c["logits"]) > 1e-2                      # dropout changed the network
    measured(f"flava_small_dropout/{precision}/logits", rel(logits.detach().cpu(), ref_logits))
    assert rel(logits.detach().cpu(), ref_logits) < tl
    assert abs(float(loss) - float(ref_loss)) < max(tl, 2e-3) * max(1.0, abs(float(ref_loss)))
    if precision == "fp32":
        assert torch.equal(logits.detach().cpu().argmax(-1), ref_logits.argmax(-1))
    for k, prm in m.named_parameters():
        g = ref_grads[k]
        scale = float(g.abs().max())
        if scale < 1e-7:
            assert float(prm.grad.abs().max()) < 1e-7, k
        else:
            measured(f"flava_small_dropout/{precision}/grad", float((prm.grad.cpu() - g).abs().max()) / scale)
            assert float((prm.grad.cpu() - g).abs().max()) < tg * scale, k
    # reproducible under torch.manual_seed; a fresh draw differs
    torch.manual_seed(77)
    again = m(x).detach()
    assert m.last_dropout_seed == seed and torch.equal(again, logits.detach())
    other = m(x).detach()
    assert m.last_dropout_seed != seed and not torch.equal(other, logits.detach())
    # eval: dropout is the identity -> the dropout-free golden
    m.eval()
    with torch.no_grad():
        assert rel(m(x).cpu(), c["logits_eval"]) < tl


def _mmbt_args(cfg, precision, **kw):
    import types
    vocab = types.SimpleNamespace(stoi={"[CLS]": cfg["cls_id"], "[SEP]": cfg["sep_id"], "[PAD]": 0})
    bc = dict(vocab=cfg["vocab"], D=cfg["D"], n_head=cfg["n_head"], n_layers=cfg["n_layers"], d_ff=cfg["d_ff"],
              max_pos=cfg["max_pos"], n_types=cfg["n_types"], init_range=0.02)
    bc.update(kw.pop("bert_config", {}))
    return types.SimpleNamespace(bert_model="test", hidden_sz=cfg["D"], img_hidden_sz=cfg["d_img"],
                                 num_image_embeds=cfg["n_img"], img_embed_pool_type="avg", dropout=kw.pop("dropout", 0.0),
                                 n_classes=cfg["C"], vocab=vocab, precision=precision, img_encoder=None,
                                 bert_config=bc, **kw)


@pytest.mark.parametrize("precision,tl,tg", [("fp32", 1e-3, 1e-3), ("bf16", 1.6e-2, 6e-2)])
@pytest.mark.parametrize("indices,s_txt", [(None, 150), ("control", 150), (None, 330)])
def test_mmbt_training_with_dropout_matches_oracle_under_the_same_masks(mmu, precision, tl, tg, indices, s_txt,
                                                                        measured):
    """MMBT in train() mode with the reference's dropouts switched on (BERT hidden 0.1, attention
    probabilities 0.15, ImageBertEmbeddings 0.2 -- src/mmbt.py:56,82 and pytorch_pretrained_bert's
    BertEmbeddings / BertSelfAttention / BertSelfOutput / BertOutput): forward + backward of the
    engine (bf16: the fused attention kernels -- undropped P stored, dropped P into P V, mask
    regenerated in the fused backward --, fp32: the SIMT path with a dropped-probability copy)
    against the oracle applying the same counter-based masks; ragged batch, S = 155 (two query
    tiles, three 64-key chunks) and S = 335 (three tiles, chunks in both column halves), also
    through a ``forward_control``-style index list."""
    from oracle import mmbt as O
    cfg = dict(B=3, S_txt=s_txt, n_img=3, d_img=64, D=128, n_head=2, n_layers=2, d_ff=256, vocab=300,
               max_pos=s_txt + 10, n_types=2, C=2, cls_id=5, sep_id=6)
    drop = dict(hidden=0.1, attn=0.15, img=0.2)
    g = torch.Generator().manual_seed(5)
    args = _mmbt_args(cfg, precision, dropout=drop["img"],
                      bert_config=dict(hidden_dropout_prob=drop["hidden"], attention_probs_dropout_prob=drop["attn"]))
    m = mmu.MultimodalBertClf(args)
    assert (m.drop_hidden, m.drop_attn, m.drop_img) == pytest.approx((0.1, 0.15, 0.2))
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.08)
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    txt = torch.randint(7, cfg["vocab"], (3, s_txt), generator=g)
    lens = torch.tensor([s_txt, (2 * s_txt) // 3 - 3, 31])
    mask = (torch.arange(s_txt)[None] < lens[:, None]).long()
    txt, segment = txt * mask, mask.clone()
    tok = torch.randn(3, 3, 64, generator=g)
    y = torch.tensor([0, 1, 1])
    idx = None
    if indices == "control":
        idx = O.control_indices(s_txt + 5, 60, torch.Generator().manual_seed(9))
    m.cuda().train()
    m.zero_grad()
    t = tok.cuda().requires_grad_(True)
    torch.manual_seed(123)
    if idx is None:
        logits = m(txt.cuda(), mask.cuda(), segment.cuda(), t)
    else:
        logits = m.forward_indices(txt.cuda(), mask.cuda(), segment.cuda(), t, [int(i) for i in idx])
    seed = m.last_dropout_seed
    loss = m.compute_loss(logits, y.cuda())
    loss.backward()
    P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}
    ref_logits, ref_loss, ref_grads, ref_dtok = O.loss_and_grads(
        P64, txt, mask, segment, tok.double(), y, cfg, idx, dropout=dict(seed=seed, **drop))
    plain_logits, _, _, _ = O.loss_and_grads(P64, txt, mask, segment, tok.double(), y, cfg, idx)
    assert rel(ref_logits, plain_logits) > 1e-2                     # dropout changed the network
    measured(f"mmbt_dropout/{precision}/logits", rel(logits.detach().cpu(), ref_logits))
    measured(f"mmbt_dropout/{precision}/dimg", rel(t.grad.cpu(), ref_dtok))
    assert rel(logits.detach().cpu(), ref_logits) < tl
    assert abs(float(loss.detach()) - float(ref_loss)) < max(tl, 2e-3) * abs(float(ref_loss))
    assert rel(t.grad.cpu(), ref_dtok) < tg
    gmax = max(float(v.abs().max()) for v in ref_grads.values())
    for k, prm in m.named_parameters():
        if float(ref_grads[k].abs().max()) < 1e-5 * gmax:
            continue
        measured(f"mmbt_dropout/{precision}/grad", rel(prm.grad.cpu(), ref_grads[k]))
        assert rel(prm.grad.cpu(), ref_grads[k]) < tg, (k, rel(prm.grad.cpu(), ref_grads[k]))
    # reproducible under torch.manual_seed; eval mode = the dropout-free network
    torch.manual_seed(123)
    again = (m(txt.cuda(), mask.cuda(), segment.cuda(), tok.cuda()) if idx is None else
             m.forward_indices(txt.cuda(), mask.cuda(), segment.cuda(), tok.cuda(), [int(i) for i in idx]))
    assert m.last_dropout_seed == seed and torch.equal(again.detach(), logits.detach())
    m.eval()
    with torch.no_grad():
        ev = (m(txt.cuda(), mask.cuda(), segment.cuda(), tok.cuda()) if idx is None else
              m.forward_indices(txt.cuda(), mask.cuda(), segment.cuda(), tok.cuda(), [int(i) for i in idx]))
    assert rel(ev.cpu(), plain_logits) < tl


def test_mmbt_default_configuration_applies_the_reference_dropout(mmu):
    """The reference's default MMBT set-up (bert-base config: hidden / attention dropout 0.1) must
    not silently train the dropout-free network: defaults are 0.1, ``bert_dropout=0`` opts out."""
    cfg = dict(B=2, S_txt=6, n_img=2, d_img=16, D=64, n_head=1, n_layers=1, d_ff=64, vocab=30, max_pos=16,
               n_types=2, C=2, cls_id=1, sep_id=2)
    m = mmu.MultimodalBertClf(_mmbt_args(cfg, "fp32"))
    assert m.drop_hidden == pytest.approx(0.1) and m.drop_attn == pytest.approx(0.1) and m.drop_img == 0.0
    m0 = mmu.MultimodalBertClf(_mmbt_args(cfg, "fp32", bert_dropout=0.0))
    assert m0.drop_hidden == 0.0 and m0.drop_attn == 0.0
    m0.load_state_dict(m.state_dict())
    g = torch.Generator().manual_seed(2)
    txt = torch.randint(3, 30, (2, 6), generator=g).cuda()
    mask = torch.ones(2, 6, dtype=torch.long).cuda()
    tok = torch.randn(2, 2, 16, generator=g).cuda()
    m.cuda().train(); m0.cuda().train()
    a, b = m(txt, mask, mask, tok).detach(), m0(txt, mask, mask, tok).detach()
    assert not torch.equal(a, b) and torch.isfinite(a).all()
    m.eval(); m0.eval()
    with torch.no_grad():
        assert torch.equal(m(txt, mask, mask, tok), m0(txt, mask, mask, tok))


def test_cls_variant_default_dropout_trains(mmu):
    """FlavaFusionTransfomerwithCLSToken defaults to drop = 0.1 (reference src/model.py:306-318):
    the default-constructed model must train."""
    torch.manual_seed(0)
    m = mmu.FlavaFusionTransfomerwithCLSToken(out_dim=2, num_classes=5, image_hidden_size=32, text_hidden_size=32,
                                              multimodal_hidden_size=64, multimodal_num_attention_heads=2,
                                              multimodal_num_hidden_layers=2, avg_pool=False).cuda().train()
    assert m.drop == pytest.approx(0.1)
    opt = mmu.FusedAdamW(m.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(1)
    img, txt = torch.randn(16, 6, 32, generator=g).cuda(), torch.randn(16, 4, 32, generator=g).cuda()
    y = torch.randint(0, 5, (16, 1), generator=g).repeat(1, 2).cuda()
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = m.compute_loss(m((img, txt)), y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(math.isfinite(v) for v in losses) and min(losses[4:]) < losses[0]
