"""Pin the CPU oracle to the reference: every golden fixture was produced by the unmodified
reference modules (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import fusion, optim, shaping, uncertainty

SMALL = ["plain_E2", "plain_E1", "avgpool_E2", "cls_E3", "plain_E5_h3"]


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", SMALL)
def test_flava_forward_loss_grads(golden, name):
    c = golden("flava_small.pt")[name]
    cfg = c["cfg"]
    P = c["state_dict"]
    logits, loss, grads = fusion.loss_and_grads(P, (c["img"], c["txt"]), c["y_train"],
                                                cfg["heads"], cfg["avg_pool"])
    assert rel_err(logits, c["logits"]) < 1e-5
    assert abs(float(loss) - float(c["loss"])) < 1e-5 * max(1.0, abs(float(c["loss"])))
    for k, g in c["grads"].items():
        scale = max(float(g.abs().max()), 1e-6)
        assert float((grads[k] - g).abs().max()) <= 2e-4 * scale + 1e-7, k
    # eval path: CE on the head-mean logits, argmax of the same
    le = fusion.flava_fusion_forward(P, (c["img"], c["txt"]), cfg["heads"], cfg["avg_pool"])
    assert rel_err(le, c["logits_eval"]) < 1e-5
    assert abs(float(fusion.compute_loss(le, c["y"], eval=True)) - float(c["loss_eval"])) < 1e-5
    assert float(fusion.acc(le, c["y"], True, True)) == float(c["eval_acc"])
    assert float(fusion.acc(logits, c["y_train"], False, True)) == float(c["train_acc"])


@pytest.mark.parametrize("name", ["plain_E2", "avgpool_E2", "plain_E5_h3"])
def test_eager_incumbent_matches_reference_goldens(golden, name):
    """oracle/eager.py -- the fused-ATen form that bench.py times on `cuda` as the PyTorch-eager
    incumbent -- computes the reference's function: logits, loss, every gradient and one
    torch.optim.AdamW step against the goldens of the unmodified reference."""
    from oracle import eager
    c = golden("flava_small.pt")[name]
    cfg = c["cfg"]
    tr = eager.EagerTrainer(c["state_dict"], cfg["heads"], cfg["layers"], cfg["E"], lr=1e-3, wd=1e-3)
    logits = eager.forward(tr.P, (c["img"], c["txt"]), cfg["heads"], cfg["layers"], cfg["E"], cfg["avg_pool"])
    assert rel_err(logits.detach(), c["logits"]) < 1e-5
    loss = eager.compute_loss(logits, c["y_train"])
    assert abs(float(loss) - float(c["loss"])) < 1e-5 * max(1.0, abs(float(c["loss"])))
    loss.backward()
    for k, g in c["grads"].items():
        got = tr.P[k].grad if tr.P[k].grad is not None else torch.zeros_like(g)
        assert float((got - g).abs().max()) <= 2e-4 * max(float(g.abs().max()), 1e-6) + 1e-7, k
    if not cfg["avg_pool"]:
        tr2 = eager.EagerTrainer(c["state_dict"], cfg["heads"], cfg["layers"], cfg["E"], lr=1e-3, wd=1e-3)
        tr2.train_step((c["img"], c["txt"]), c["y_train"])
        for k, g in c["grads"].items():
            ok = g.abs() > 1e-3 * g.abs().max().clamp_min(1e-12)
            d = (tr2.P[k].detach() - c["params_after_adamw"][k]).abs()
            assert float(d[ok].max() if ok.any() else 0.0) < 2e-5, k


def test_flava_missing_modality(golden):
    c = golden("flava_small.pt")["cls_E3"]
    P, h = c["state_dict"], c["cfg"]["heads"]
    assert rel_err(fusion.flava_fusion_forward(P, (c["img"], None), h), c["logits_img_only"]) < 1e-5
    assert rel_err(fusion.flava_fusion_forward(P, (None, c["txt"]), h), c["logits_txt_only"]) < 1e-5


def test_dead_text_projection_grad(golden):
    """without avg_pool only positions < E are live."""
    c = golden("flava_small.pt")["plain_E2"]
    assert float(c["grads"]["text_to_mm_projection.weight"].abs().sum()) == 0.0


@pytest.mark.parametrize("name", SMALL)
def test_adamw_one_step(golden, name):
    c = golden("flava_small.pt")[name]
    for k, g in c["grads"].items():
        p0 = c["state_dict"][k]
        p1, _, _ = optim.adamw_step(p0, g, torch.zeros_like(p0), torch.zeros_like(p0), 1, 1e-3)
        assert float((p1 - c["params_after_adamw"][k]).abs().max()) < 2e-6, k


def test_adamw_cosine_trajectory(golden):
    c = golden("adamw_cosine.pt")
    p = c["p0"].clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for t, g in enumerate(c["grads"]):
        lr = c["lr"] * optim.cosine_with_warmup_factor(t, c["warmup"], c["total"])
        assert abs(lr - c["lrs"][t]) < 1e-12
        p, m, v = optim.adamw_step(p, g, m, v, t + 1, lr)
        assert float((p - c["traj"][t]).abs().max()) < 5e-6
    assert float((m - c["exp_avg"]).abs().max()) < 1e-6
    assert float((v - c["exp_avg_sq"]).abs().max()) < 1e-6


def test_full_width_model(golden):
    from tests.golden.make_golden import det_state_dict
    c = golden("flava_768.pt")
    P = det_state_dict(c["shapes"], c["cfg"]["seed"])
    logits, loss, grads = fusion.loss_and_grads(P, (c["img"], c["txt"]), c["y_train"], 3, False)
    assert rel_err(logits, c["logits"]) < 2e-5
    assert abs(float(loss) - float(c["loss"])) < 1e-5 * float(c["loss"])
    for k, s in c["grad_summaries"].items():
        g = grads[k].double()
        got = torch.stack([g.sum(), g.abs().sum(), g.pow(2).sum()])
        assert torch.allclose(got[1:], s[1:], rtol=2e-3, atol=1e-7), k
        assert abs(float(got[0] - s[0])) <= 1e-5 * float(s[1]) + 1e-7, k  # signed sum cancels
    for k, sl in c["grad_slices"].items():
        scale = max(float(sl.abs().max()), 1e-6)
        assert float((grads[k].reshape(-1)[:64] - sl).abs().max()) < 1e-3 * scale + 1e-7, k


def test_mimo_transformer(golden):
    c = golden("mimo_transformer.pt")
    logits, loss, grads = fusion.loss_and_grads(c["state_dict"], c["x"], c["y_train"],
                                                c["cfg"]["heads"], model="mimo")
    assert rel_err(logits, c["logits"]) < 1e-5
    assert abs(float(loss) - float(c["loss"])) < 1e-5
    for k, g in c["grads"].items():
        scale = max(float(g.abs().max()), 1e-6)
        assert float((grads[k] - g).abs().max()) <= 2e-4 * scale + 1e-7, k


def test_shaping_bit_exact(golden):
    c = golden("shaping.pt")
    inp = c["inputs"]
    for mt in ["Vanilla", "MultiHead", "MIMO-shuffle-instance"]:
        for phase in ["train", "eval"]:
            torch.manual_seed(42)
            (i2, t2), y2 = shaping.data_forming_func_transformer((inp["img"], inp["txt"]),
                                                                 inp["y"], phase, mt)
            g = c[f"transformer/{mt}/{phase}"]
            assert torch.equal(i2, g["img"]) and torch.equal(t2, g["txt"]) and torch.equal(y2, g["y"])
    for mt in ["Vanilla", "single-model-weight-sharing", "MultiHead", "MIMO-shuffle-instance",
               "MIMO-shuffle-view", "MIMO-shuffle-all"]:
        for phase in ["train", "eval"]:
            torch.manual_seed(42)
            x2, y2 = shaping.data_forming_func(inp["x"], inp["yv"], phase, mt)
            g = c[f"fmnist/{mt}/{phase}"]
            assert torch.equal(x2, g["x"]) and torch.equal(y2, g["y"]), (mt, phase)
    col = c["collate"]
    (pi, pt), pl = shaping.collate_fn_flava(col["ragged"])
    assert torch.equal(pi, col["img"]) and torch.equal(pt, col["txt"]) and torch.equal(pl, col["labels"])


def test_input_sampling_bit_exact(golden):
    c = golden("input_sampling.pt")
    np.random.seed(c["np_seed"])
    torch.manual_seed(c["torch_seed"])
    variants = shaping.robustness_variants(c["l_img"], c["l_txt"], c["n_repeats"])
    assert len(variants) == 3 + 2 * c["n_repeats"]
    for (ii, it), (gi, gt) in zip(variants[3:], c["draws"]):
        ii = ii if ii is not None else torch.zeros(0, dtype=torch.int64)
        it = it if it is not None else torch.zeros(0, dtype=torch.int64)
        assert torch.equal(ii, gi) and torch.equal(it, gt)


def test_notebook_scoring(golden):
    c = golden("notebook_scoring.pt")
    preds, labels = c["preds"].numpy(), c["labels"].numpy()
    ori, image, text, ic, tc = uncertainty.process_predictions(preds, labels)
    assert np.allclose(ori, c["ori"].numpy(), rtol=1e-6)
    assert np.allclose(ic, c["image_corr"].numpy(), rtol=1e-6)
    assert np.allclose(tc, c["text_corr"].numpy(), rtol=1e-6)
    corr = uncertainty.get_correlation(ori, image, text, ic, tc)
    assert abs(corr["image"] - c["corr_image"]) < 1e-6  # reference runs pearsonr in fp32
    assert abs(corr["text"] - c["corr_text"]) < 1e-6
    assert abs(uncertainty.acc_table(preds, labels)["full"] - c["acc_full"]) < 1e-9


def test_uncertainty_identities():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(500, 5, 101, generator=g) * 3
    y = torch.randint(0, 101, (500,), generator=g)
    s = uncertainty.ensemble_scores(logits)
    assert float(s["mi"].min()) > -1e-12                      # MI >= 0 (Jensen)
    assert float((s["h_pred"] - np.log(101)).max()) < 1e-12    # H <= log C
    same = logits[:, :1].expand(-1, 5, -1)                     # identical heads -> MI = 0
    assert float(uncertainty.ensemble_scores(same)["mi"].abs().max()) < 1e-12
    h = uncertainty.calibration_histograms(logits, y)
    assert int(h["conf_count"].sum()) == 500 and int(h["hpred_count"].sum()) == 500
    assert int(h["conf_correct"].sum()) == h["n_correct_prob"]
    ece = uncertainty.ece_from_bins(h["conf_count"], h["conf_correct"], h["conf_sum"])
    assert 0.0 <= ece <= 1.0


def test_mimo_resnet(golden):
    """oracle/resnet.py against the reference's MIMOResNet (src/model.py:17-100, layers.py:7-38):
    train-mode logits, loss, every gradient, BatchNorm running statistics, eval-mode logits.
    The golden's seed was searched (tests/golden/make_golden.py) so that no ReLU pre-activation
    lies within 2e-5 of zero: a pre-activation within fp32 rounding of zero flips its mask in one
    implementation and not the other and, with 4 samples, moves BatchNorm gradients by percents."""
    from oracle import resnet
    rel = rel_err
    c = golden("mimo_resnet.pt")
    C = c["cfg"]["C"]
    P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in c["state_dict"].items()}
    logits, loss, grads, buffers = resnet.loss_and_grads(P64, c["x"].double(), c["y_train"], C)
    assert rel(logits, c["logits"]) < 1e-5 and abs(float(loss) - float(c["loss"])) < 1e-5
    assert torch.equal(logits.argmax(-1), c["logits"].argmax(-1))
    for k, g in c["grads"].items():
        assert rel(grads[k], g) < 5e-5, k
    for k, v in buffers.items():
        assert rel(v.double(), c["state_after_forward"][k].double()) < 1e-5, k
    after64 = {k: (v.double() if v.is_floating_point() else v) for k, v in c["state_after_forward"].items()}
    ev = resnet.mimo_resnet_forward(after64, c["x"].double(), C, training=False)
    assert rel(ev, c["logits_eval"]) < 1e-5
    # fp32 oracle (the dtype the CUDA engine is compared in)
    l32, loss32, g32, _ = resnet.loss_and_grads(c["state_dict"], c["x"], c["y_train"], C)
    assert rel(l32, c["logits"]) < 1e-5
    for k, g in c["grads"].items():
        assert rel(g32[k], g) < 2e-4, k


# ------------------------------------------------------------------------------- MMBT path
@pytest.mark.parametrize("name", ["fp32_small", "hd64"])
def test_mmbt_oracle_matches_reference(golden, name):
    """oracle/mmbt.py against the UNMODIFIED reference src/mmbt.py (run on the restated third-party
    BERT, tests/golden/make_golden_mmbt.py): all four forward entry points, the index draw of
    forward_control, loss, every gradient and the gradient of the pooled image tokens."""
    from oracle import mmbt as O
    c = golden("mmbt_small.pt")[name]
    cfg = c["cfg"]
    P = {k: (v.double() if v.is_floating_point() else v) for k, v in c["state_dict"].items()}
    tok = c["img_tokens"].double()
    x = (c["txt"], c["mask"], c["segment"], tok, cfg)
    assert rel_err(O.forward(P, *x), c["logits_full"]) < 1e-5
    for mode in ("img_only", "txt_only"):
        idx = O.mode_indices(mode, cfg["n_img"], cfg["S_txt"])
        assert rel_err(O.forward(P, *x, idx), c["logits_" + mode]) < 1e-5, mode
    for modal, d in c["control"].items():
        torch.manual_seed(d["seed"])
        num = cfg["n_img"] + 1 if modal == "image" else cfg["S_txt"]
        ind = O.control_indices(cfg["S_txt"] + cfg["n_img"] + 2, num)
        assert torch.equal(ind, d["indices"])  # bit-exact index draw
        assert rel_err(O.forward(P, *x, ind), d["logits"]) < 1e-5, modal
    logits, loss, grads, dtok = O.loss_and_grads(P, c["txt"], c["mask"], c["segment"], tok, c["y"], cfg)
    assert abs(float(loss) - float(c["loss"])) < 1e-6
    assert rel_err(dtok, c["dimg_tokens"]) < 1e-4
    gmax = max(float(g.abs().max()) for g in c["grads"].values())
    for k, g in c["grads"].items():
        if float(g.abs().max()) < 1e-5 * gmax:  # analytically zero (key bias); fp32 rounding noise
            assert float(grads[k].abs().max()) < 1e-5 * gmax, k
        else:
            assert rel_err(grads[k], g) < 2e-4, k


def test_bertadam_oracle_matches_restated_reference_optimizer(golden):
    from oracle import mmbt as O
    c = golden("bertadam.pt")
    h = c["hyper"]
    st = {k: dict(p=v.double(), m=torch.zeros_like(v).double(), v=torch.zeros_like(v).double())
          for k, v in c["init"].items()}
    for step, (gr, after) in enumerate(zip(c["grads"], c["after"])):
        for k, s in st.items():
            s["p"], s["m"], s["v"] = O.bertadam_step(
                s["p"], gr[k].double(), s["m"], s["v"], step, lr=h["lr"], warmup=h["warmup"],
                t_total=h["t_total"], weight_decay=c["decay"][k], b1=h["b1"], b2=h["b2"], e=h["e"],
                max_grad_norm=h["max_grad_norm"])
            assert rel_err(s["p"], after[k]) < 1e-5, (step, k)


@pytest.mark.parametrize("name", ["avg3", "max4"])
def test_image_encoder_oracle_matches_reference(golden, name):
    """oracle/image_encoder.py against the reference's unmodified ImageEncoder class
    (src/mmbt.py:15-45) over a thin torchvision Bottleneck ResNet: eval tokens, train-mode tokens,
    running statistics and every gradient (large tensors through a strided digest).  The seeds
    were chosen so that no ReLU / max-pool decision sits within fp32 rounding of a tie."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from det_params import det_image_encoder_state, digest_error
    from oracle import image_encoder as IE
    c = golden("image_encoder.pt")[name]
    cfg = c["cfg"]
    sd = det_image_encoder_state(c["state_dict_shapes"], cfg["seed"])
    P = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    pool = {3: (3, 1), 4: (2, 2)}[cfg["n_img"]]
    is_max = cfg["pool"] != "avg"
    te = IE.image_encoder_forward(P, c["x"].double(), cfg["layers"], pool, is_max, False)
    assert rel_err(te, c["tokens_eval"]) < 1e-5
    tt, grads, buf = IE.tokens_and_grads(P, c["x"].double(), c["r"].double(), cfg["layers"], pool, is_max)
    assert rel_err(tt, c["tokens_train"]) < 5e-5
    for k, d in c["grads"].items():
        assert digest_error(d, grads[k]) < 1e-4, k
    for k, v in c["buffers_after"].items():
        assert rel_err(buf[k], v) < 1e-5, k


def test_rank_stats_against_reference_notebooks(golden):
    """oracle/rank_stats.py against the reference's own trunk_pred_top / subnetwork_wise_kendalltau
    (notebooks/analysis_round_1.py:74-90) and AUC_table (notebooks/hatefulmeme_robustness.py:22-41),
    executed unmodified by tests/golden/make_golden.py::rank_case."""
    from oracle import rank_stats
    c = golden("rank_stats.pt")
    preds, labels = c["predictions"].numpy(), c["labels"].numpy()
    for k in range(preds.shape[1]):
        got = rank_stats.trunk_pred_top(preds[:, k], labels, c["top"], mute_true=True)
        assert np.array_equal(got.astype(np.float64), c["muted"][k].numpy())      # bit-exact
    got = rank_stats.trunk_pred_top(preds[:, 0], labels, 3, mute_true=False)
    assert np.array_equal(got.astype(np.float64), c["plain_top3"].numpy())
    taus = rank_stats.subnetwork_wise_kendalltau(preds, labels, c["top"])
    assert np.allclose(taus, c["taus"].numpy(), rtol=0, atol=1e-12)
    auc = rank_stats.auc_table(c["hm_labels"].numpy(), c["hm_scores"].numpy())
    assert np.allclose(auc, c["hm_auc"].numpy(), rtol=0, atol=1e-12)
    # the scores themselves: head-mean probability of class 1 (process_predictions_hatefulmeme)
    p1 = uncertainty.notebook_softmax(c["hm_preds"].numpy()).mean(2)[..., 1]
    assert np.allclose(p1, c["hm_scores"].numpy(), rtol=1e-6)


def test_rank_stats_against_scipy_and_sklearn():
    """The third-party definitions the reference calls, live, on vectors with heavy ties."""
    import scipy.stats as stats
    from sklearn.metrics import roc_auc_score
    from oracle import rank_stats
    rng = np.random.RandomState(4)
    for n, levels in [(2, 2), (3, 50), (257, 4), (1500, 1000), (2500, 7)]:
        x = rng.randint(0, levels, size=n).astype(np.float32)
        y = rng.randint(0, levels, size=n).astype(np.float32) + (x > levels // 2)
        if len(set(x)) > 1 and len(set(y)) > 1:
            assert abs(rank_stats.kendalltau(x, y) - stats.kendalltau(x, y)[0]) < 1e-12
        lab = rng.randint(0, 2, size=n)
        if 0 < lab.sum() < n:
            assert abs(rank_stats.auroc(lab, y) - roc_auc_score(lab, y)) < 1e-12
    assert np.isnan(rank_stats.kendalltau(np.ones(5), np.arange(5)))
    conc, disc, tx, ty = rank_stats.pair_counts([1, 2, 2, 3], [1, 3, 3, 2])
    assert (conc, disc, tx, ty) == (3, 2, 1, 1)


def test_rank_stats_count_identities():
    """Size-independent properties of the pair counts the device kernel is held to: the five
    classes partition the n(n-1)/2 pairs, swapping the arguments swaps the tie counts, negating one
    argument swaps concordant and discordant, a strictly monotone map changes nothing."""
    from oracle import rank_stats
    rng = np.random.RandomState(11)
    for n in (1, 2, 5, 64, 700):
        x = rng.randint(0, 5, size=n).astype(np.float32)
        y = rng.randint(0, 9, size=n).astype(np.float32)
        c, d, tx, ty = rank_stats.pair_counts(x, y)
        tot = n * (n - 1) // 2
        joint = c + d + tx + ty - tot
        assert 0 <= joint <= min(tx, ty)
        ux, cx = np.unique(x, return_counts=True)
        assert tx == int((cx * (cx - 1) // 2).sum())                     # ties in x from multiplicities
        assert rank_stats.pair_counts(y, x) == (c, d, ty, tx)
        assert rank_stats.pair_counts(x, -y) == (d, c, tx, ty)
        assert rank_stats.pair_counts(3 * x + 1, np.exp(y / 4)) == (c, d, tx, ty)
    # AUROC is 1/2 for constant scores and 1 for a perfect ranking
    lab = np.array([0, 1, 1, 0, 1])
    assert rank_stats.auroc(lab, np.zeros(5)) == 0.5
    assert rank_stats.auroc(lab, lab + 0.0) == 1.0
    assert np.isnan(rank_stats.auroc(np.ones(4), np.arange(4.0)))        # one class only


def test_restated_bert_matches_transformers_bertmodel():
    """oracle/bert_restated.py (the absent ``pytorch_pretrained_bert`` 0.6.2 restated) against the
    maintained successor of that package, ``transformers.BertModel`` (installed: 5.5.0): same
    state-dict keys (strict load), same embeddings / encoder / pooler arithmetic to 1e-12 in fp64
    on a padded batch.  This pins the BERT layer arithmetic the MMBT goldens are built on to an
    independent implementation; the call sequence on the restated side is the reference's own
    (src/mmbt.py:103-128: additive mask (1 - m) * -10000, ``bert.encoder(..., output_all_encoded_layers
    =False)``, ``bert.pooler``)."""
    tf = pytest.importorskip("transformers")
    from oracle import bert_restated as R
    kw = dict(vocab_size=97, hidden_size=48, num_hidden_layers=3, num_attention_heads=4,
              intermediate_size=80, max_position_embeddings=40)
    torch.manual_seed(0)
    m = R.BertModel(R.BertConfig(**kw)).double().eval()
    for p in m.parameters():          # move biases / LayerNorm parameters off their trivial initial values
        p.data.add_(torch.randn_like(p) * 0.05)
    hf = tf.BertModel(tf.BertConfig(hidden_act="gelu", layer_norm_eps=1e-12, **kw)).double().eval()
    missing, unexpected = hf.load_state_dict(m.state_dict(), strict=False)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 97, (3, 17), generator=g)
    ids[1, 12:] = 0
    tt = torch.randint(0, 2, (3, 17), generator=g)
    mask = (ids != 0).long()
    with torch.no_grad():
        o = hf(input_ids=ids, token_type_ids=tt, attention_mask=mask)
        ext = (1.0 - mask[:, None, None, :].double()) * -10000.0
        last = m.encoder(m.embeddings(ids, tt), ext, output_all_encoded_layers=False)[-1]
        pooled = m.pooler(last)
    keep = mask.bool()
    assert float((o.last_hidden_state - last)[keep].abs().max()) < 1e-12
    assert float((o.pooler_output - pooled).abs().max()) < 1e-12


def test_fmnist_view_format_and_view_sweep(golden):
    """Oracle restatements of the FashionMNIST view handling against the reference run in the build
    container (tests/golden/make_golden.py::fmnist_views_case): ``quarter_views`` against the
    reference's QuarterCrop + ToTensor composition on uint8 images (bit-exact), and the sweep
    inputs (zero-filled view / removed view) against the reference script's own sweep statements
    executed over a linear stand-in model (both model types, outputs and saved labels)."""
    from oracle import shaping
    c = golden("fmnist_views.pt")
    imgs = c["images_u8"].float().div(255).unsqueeze(1)                  # ToTensor of an 'L' image
    q = torch.stack([shaping.quarter_views(im) for im in imgs])          # per sample, as the transform does
    assert torch.equal(q, c["quarters"])
    C_ = 10
    for mt, fwd in (("MultiHead", lambda x: (x.reshape(x.shape[0], -1) @ c["W4"]).view(-1, 4, C_)),
                    ("single-model-weight-sharing", lambda x: (x.reshape(x.shape[0], -1) @ c["W1"]).view(-1, 1, C_))):
        outs, labels = [], []
        for i in range(4):
            per = []
            for x, y in c["valid"]:
                if mt == "MultiHead":
                    per.append(fwd(shaping.leave_one_view_out(x, i)))
                    y_saved = y
                else:
                    x_, y_saved = shaping.data_forming_func(shaping.drop_one_view(x, i), y, "eval", mt)
                    per.append(fwd(x_).view(x.shape[0], 3, C_))
                if i == 0:
                    labels.append(y_saved)
            outs.append(torch.cat(per))
        assert torch.equal(torch.stack(outs), c[mt]["outputs"])
        assert torch.equal(torch.cat(labels), c[mt]["labels"])


def test_mmbt_collate_bit_exact(golden):
    """oracle.shaping.collate_fn against the reference's own MMBT ``collate_fn`` (src/dataset.py:420-438)."""
    from oracle import shaping
    c = golden("mmbt_collate.pt")
    (txt, segment, mask, img), tgt = shaping.collate_fn(c["rows"])
    for got, key in ((txt, "txt"), (segment, "segment"), (mask, "mask"), (img, "img"), (tgt, "tgt")):
        assert got.dtype == c[key].dtype and torch.equal(got, c[key]), key
