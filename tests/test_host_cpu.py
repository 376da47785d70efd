"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every declared symbol,
the Python boundary mirrors the reference (keys, parameter order, seeded init, shaping, index
sampling, LR schedule), fails loudly without a GPU, and the N>1 helpers work over gloo."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mmu():
    import mmu_b200
    return mmu_b200


def test_library_exports_every_declared_symbol(mmu):
    header = open(os.path.join(ROOT, "include", "mmu_b200.h")).read()
    declared = re.findall(r"MMU_API[^;(]*?\b(mmu_\w+)\s*\(", header)
    assert len(declared) >= 17 and len(set(declared)) == len(declared)
    for name in declared:
        assert hasattr(mmu._lib.lib, name), name
    assert sorted(declared) == sorted(mmu._lib.EXPORTS)
    assert b"sm_100a" in mmu._lib.lib.mmu_version()
    assert mmu._lib.lib.mmu_error_string(-7) == b"workspace too small"


def test_struct_sizes_match_header(mmu):
    import ctypes as C
    assert C.sizeof(mmu._lib.MetricAccum) == (15 + 15 + 32 + 32 + 4) * 8 + (15 + 4) * 8
    assert C.sizeof(mmu._lib.FlavaConfig) == 15 * 4
    # the ctypes mirrors against sizeof() as compiled into the library
    for which, klass in enumerate((mmu._lib.FlavaConfig, mmu._lib.FlavaInputs, mmu._lib.GemmEpilogue,
                                   mmu._lib.MetricAccum, mmu._lib.ParamEntry,
                                   mmu._lib.PosthocAccum)):
        assert mmu._lib.lib.mmu_struct_size(which) == C.sizeof(klass), klass.__name__
    assert mmu._lib.lib.mmu_struct_size(6) == C.sizeof(mmu._lib.MmbtConfig) == 20 * 4
    assert mmu._lib.lib.mmu_struct_size(7) == C.sizeof(mmu._lib.MmbtInputs)
    assert C.sizeof(mmu._lib.ParamEntry) == 96 + 8 + 8 + 12 + 4  # padded to 8
    assert mmu._lib.ACC_OFF["conf_sum"] == 98


def small_model(mmu, cls=False, **kw):
    klass = mmu.FlavaFusionTransfomerwithCLSToken if cls else mmu.FlavaFusionTransfomer
    args = dict(out_dim=2, num_classes=7, image_hidden_size=32, text_hidden_size=48,
                multimodal_hidden_size=64, multimodal_num_attention_heads=2,
                multimodal_num_hidden_layers=2, drop=0.0, avg_pool=False)
    args.update(kw)
    return klass(**args)


def test_state_dict_keys_and_strict_load(mmu, golden):
    c = golden("flava_small.pt")["plain_E2"]
    m = small_model(mmu)
    assert set(m.state_dict()) == set(c["state_dict"])
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(c["state_dict"][k].shape), k
    m.load_state_dict(c["state_dict"], strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, c["state_dict"][k])
    # parameters are views of one flat buffer, gradients of one flat gradient buffer
    base = m._flat.data_ptr()
    for p in m.parameters():
        assert base <= p.data_ptr() < base + m._flat.numel() * 4
        assert p.grad is not None and p.grad.shape == p.shape
    c3 = golden("flava_small.pt")["cls_E3"]
    m3 = small_model(mmu, cls=True, out_dim=3)
    m3.load_state_dict(c3["state_dict"], strict=True)


def test_seeded_init_and_parameter_order_match_reference(mmu, golden):
    g = golden("init_seed123.pt")
    for name, cls, kw in [("plain", False, dict(out_dim=2, num_classes=7)),
                          ("cls", True, dict(out_dim=3, num_classes=5))]:
        torch.manual_seed(123)
        m = small_model(mmu, cls=cls, **kw)
        assert [k for k, _ in m.named_parameters()] == g[name + "_param_order"]
        for k, v in m.state_dict().items():
            got = torch.stack([v.double().sum(), v.double().abs().sum()])
            assert torch.allclose(got, g[name][k], rtol=1e-12, atol=1e-12), k


def test_mimo_transformer_seeded_init_and_keys(mmu, golden):
    """MIMOTransfomer: reference parameter order / state-dict keys and seed-for-seed initial
    weights (reference src/model.py:114-137)."""
    g = golden("init_seed123.pt")
    torch.manual_seed(123)
    m = mmu.MIMOTransfomer(out_dim=4, num_classes=10, hidden_size=48,
                           multimodal_num_hidden_layers=2, multimodal_num_attention_heads=2)
    assert [k for k, _ in m.named_parameters()] == g["mimo_param_order"]
    assert list(m.state_dict().keys()) == list(g["mimo"].keys())
    for k, v in m.state_dict().items():
        got = torch.stack([v.double().sum(), v.double().abs().sum()])
        assert torch.allclose(got, g["mimo"][k], rtol=1e-12, atol=1e-12), k
    with pytest.raises(mmu._lib.MMUError):   # no CPU execution path here either
        m(torch.randn(2, 4, 1, 14, 14))


def test_mimo_resnet_seeded_init_keys_and_strict_load(mmu, golden):
    """MIMOResNet: reference parameter order, state-dict keys (BatchNorm buffers included, in the
    reference's order), seed-for-seed initial weights (src/model.py:17-100), strict checkpoint load."""
    g, c = golden("init_seed123.pt"), golden("mimo_resnet.pt")
    torch.manual_seed(123)
    m = mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=4, num_classes=10)
    assert [k for k, _ in m.named_parameters()] == g["resnet_param_order"] == c["param_order"]
    assert [k for k, _ in m.named_buffers()] == c["buffer_order"]
    assert list(m.state_dict().keys()) == list(c["state_dict"].keys())
    for k, v in m.state_dict().items():
        got = torch.stack([v.double().sum(), v.double().abs().sum()])
        assert torch.allclose(got, g["resnet"][k], rtol=1e-12, atol=1e-12), k
    m.load_state_dict(c["state_dict"], strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, c["state_dict"][k]), k
    base = m._flat.data_ptr()
    for p in m.parameters():
        assert base <= p.data_ptr() < base + m._flat.numel() * 4 and p.grad is not None
    with pytest.raises(mmu._lib.MMUError):
        m(torch.rand(2, 4, 1, 14, 14))


def test_stage_ranges_partition_the_flat_buffer(mmu):
    m = small_model(mmu)
    ranges = m.stage_ranges()
    assert len(ranges) == 2 + 2  # heads, two blocks, stem
    order = sorted(ranges)
    assert order[0][0] == 0
    for (a0, a1), (b0, b1) in zip(order, order[1:]):
        assert a1 <= b0
    assert ranges[0][0] > ranges[-1][0]  # heads live at the end of the buffer, stem at the start


def test_train_eval_toggle_matches_nn_module(mmu):
    """The model's train() / eval() write `training` directly instead of walking the module tree with
    __setattr__ (a live-token step is host bound): the flags must end up exactly where
    nn.Module.train puts them, also after a submodule is added."""
    m = mmu.FlavaFusionTransfomer(out_dim=2, num_classes=5, image_hidden_size=16, text_hidden_size=16,
                                  multimodal_hidden_size=32, multimodal_num_attention_heads=2,
                                  multimodal_num_hidden_layers=2, avg_pool=False)
    assert m.eval() is m and not any(x.training for x in m.modules())
    assert m.train() is m and all(x.training for x in m.modules())
    m.add_module("extra", torch.nn.Dropout(0.5))
    m.eval()
    assert not m.extra.training and not any(x.training for x in m.modules())
    torch.nn.Module.train(m, True)          # the generic walk and the fast path agree
    assert all(x.training for x in m.modules())
    with pytest.raises(ValueError):
        m.train("yes")


def test_no_cpu_fallback(mmu):
    m = small_model(mmu)
    with pytest.raises(mmu._lib.MMUError):
        m((torch.randn(2, 3, 32), torch.randn(2, 2, 48)))
    with pytest.raises(TypeError):
        m.half()
    lin = torch.nn.Linear(4, 4)
    with pytest.raises(ValueError):
        mmu.FusedAdamW(lin.parameters())


def test_product_shaping_bit_exact(mmu, golden):
    c = golden("shaping.pt")
    inp = c["inputs"]
    ds = mmu.dataset
    for mt in ["Vanilla", "MultiHead", "MIMO-shuffle-instance"]:
        for phase in ["train", "eval"]:
            torch.manual_seed(42)
            (i2, t2), y2 = ds.data_forming_func_transformer((inp["img"], inp["txt"]), inp["y"], phase, mt)
            g = c[f"transformer/{mt}/{phase}"]
            assert torch.equal(i2, g["img"]) and torch.equal(t2, g["txt"]) and torch.equal(y2, g["y"])
    for mt in ["Vanilla", "single-model-weight-sharing", "MultiHead", "MIMO-shuffle-instance",
               "MIMO-shuffle-view", "MIMO-shuffle-all"]:
        for phase in ["train", "eval"]:
            torch.manual_seed(42)
            x2, y2 = ds.data_forming_func(inp["x"], inp["yv"], phase, mt)
            g = c[f"fmnist/{mt}/{phase}"]
            assert torch.equal(x2, g["x"]) and torch.equal(y2, g["y"]), (mt, phase)
    (pi, pt), pl = ds.collate_fn_flava(c["collate"]["ragged"])
    assert torch.equal(pi, c["collate"]["img"]) and torch.equal(pt, c["collate"]["txt"])
    assert torch.equal(pl, c["collate"]["labels"])


def test_product_index_sampling_bit_exact(mmu, golden):
    c = golden("input_sampling.pt")
    np.random.seed(c["np_seed"]); torch.manual_seed(c["torch_seed"])
    variants = mmu.robustness.robustness_variants(c["l_img"], c["l_txt"], c["n_repeats"])
    assert len(variants) == 43
    empty = torch.zeros(0, dtype=torch.int64)
    for (ii, it), (gi, gt) in zip(variants[3:], c["draws"]):
        assert torch.equal(ii if ii is not None else empty, gi)
        assert torch.equal(it if it is not None else empty, gt)
    from oracle import shaping
    for mode in ("random", "guided"):
        sc = torch.rand(64, 2, generator=torch.Generator().manual_seed(1))
        a = mmu.robustness.modality_dropout_mask(64, 0.4, mode, sc, torch.Generator().manual_seed(7))
        b = shaping.modality_dropout_mask(64, 0.4, mode, sc, torch.Generator().manual_seed(7))
        assert torch.equal(a, b) and 0 < int((a == 0).sum()) < 64
    torch.manual_seed(3); a = mmu.robustness.mask_level_variant(197, 40, "image", 4)
    torch.manual_seed(3); b = shaping.mask_level_variant(197, 40, "image", 4)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and len(a[0]) == round(4 * 197 / 9)


def test_cosine_schedule_matches_reference_lrs(mmu, golden):
    c = golden("adamw_cosine.pt")
    fn = mmu.optim.cosine_with_warmup_lambda(c["warmup"], c["total"])
    for t, lr in enumerate(c["lrs"]):
        assert abs(c["lr"] * fn(t) - lr) < 1e-12


def test_fused_adamw_state_dict_layout(mmu):
    m = small_model(mmu)
    opt = mmu.FusedAdamW(m.parameters(), lr=3e-4)
    sd = opt.state_dict()
    n = len(list(m.parameters()))
    assert sd["param_groups"][0]["params"] == list(range(n)) and len(sd["state"]) == n
    assert sd["param_groups"][0]["betas"] == (0.9, 0.98) and sd["param_groups"][0]["eps"] == 1e-9
    ref = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in m.parameters()], lr=3e-4)
    ref.load_state_dict(sd)  # same layout as torch.optim.AdamW's
    opt.load_state_dict(sd)
    m.zero_grad(); opt.zero_grad()
    assert all(p.grad is not None for p in m.parameters())  # views survive zero_grad


def test_trainer_protocol_with_a_plain_torch_model(mmu, tmp_path):
    """Model_ is model-agnostic host logic: drive it with a CPU nn.Module and torch SGD and
    check logs, callback order, history.csv / checkpoint artefacts and the stopping rule."""
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(6, 3 * 2)
        def forward(self, x):
            img, txt = x
            return self.fc(torch.cat([img.mean(1), txt.mean(1)], -1)).view(-1, 2, 3)
        def compute_loss(self, y_hat, y, eval=False):
            y_hat = y_hat.mean(1) if eval else y_hat.reshape(-1, 3)
            return torch.nn.functional.cross_entropy(y_hat, y.reshape(-1))
    def acc(y_pred, y_true, eval, dummy_dim=False):
        y_pred = y_pred.mean(1) if eval else y_pred.reshape(-1, 3)
        return (y_pred.argmax(1) == y_true.reshape(-1)).float().mean() * 100
    torch.manual_seed(0)
    from functools import partial
    train, val, test = mmu.dataset.get_synthetic_flava(4, 12, 8, 8, l_img=3, l_txt=2, dim=3, num_classes=3)
    net = Tiny()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    H, events = {}, []
    tl = mmu.src.training_loop
    cbs = tl._construct_default_callbacks(net, opt, H, str(tmp_path), checkpoint_monitor="val_acc")
    cbs.append(mmu.src.callbacks.LambdaCallback(
        on_epoch_begin=lambda e, l: events.append(("eb", e)),
        on_batch_end=lambda b, l: events.append(("be", b, set(l))),
        on_backward_end=lambda b: events.append(("bw", b))))
    for cb in cbs:
        cb.set_save_path(str(tmp_path)); cb.set_model(net, ignore=False); cb.set_optimizer(opt)
    trainer = mmu.Model_(net, opt, sched, partial(mmu.dataset.data_forming_func_transformer,
                                                  model_type="MultiHead"), metrics=[acc], verbose=False)
    trainer.to(torch.device("cpu"))
    trainer.train_loop(train, valid_generator=val, test_generator=test, epochs=2,
                       steps_per_epoch=len(train), validation_steps=len(val), test_steps=len(test),
                       callbacks=cbs, scheduler_step_on="batch", scheduler_metric=None)
    assert H["epoch"] == [1, 2] and len(H["loss"]) == 2 and "val_acc" in H and "test_loss" in H
    for f in ("history.csv", "model_epoch_1.pt", "model_last_epoch.pt", "model_best_val.pt"):
        assert os.path.exists(tmp_path / f), f
    be = [e for e in events if e[0] == "be"]
    assert be[0][2] >= {"batch", "size", "time", "batch_begin_time", "loss", "acc"}
    assert events[0] == ("eb", 1) and events[1] == ("bw", 1)


def test_train_loop_deferred_readback_gives_the_same_epoch_logs(mmu):
    """``train_loop(..., metrics_every=N)`` keeps loss / metrics on the device and reads the
    size-weighted sums back every N steps and at epoch end: same epoch means as the reference's
    per-step ``.item()`` protocol (src/framework.py:305-312), NaN stop rule kept."""
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(6, 3 * 2)
            self.poison = False
        def forward(self, x):
            img, txt = x
            return self.fc(torch.cat([img.mean(1), txt.mean(1)], -1)).view(-1, 2, 3)
        def compute_loss(self, y_hat, y, eval=False):
            y_hat = y_hat.mean(1) if eval else y_hat.reshape(-1, 3)
            loss = torch.nn.functional.cross_entropy(y_hat, y.reshape(-1))
            return loss * float("nan") if self.poison else loss
    def acc(y_pred, y_true, eval, dummy_dim=False):
        y_pred = y_pred.mean(1) if eval else y_pred.reshape(-1, 3)
        return (y_pred.argmax(1) == y_true.reshape(-1)).float().mean() * 100
    from functools import partial
    logs = {}
    for every in (1, 3, 100):
        torch.manual_seed(0)
        train, val, _ = mmu.dataset.get_synthetic_flava(4, 28, 8, 8, l_img=3, l_txt=2, dim=3, num_classes=3)
        net = Tiny()
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        trainer = mmu.Model_(net, opt, None, partial(mmu.dataset.data_forming_func_transformer,
                                                     model_type="MultiHead"), metrics=[acc], verbose=False)
        trainer.to(torch.device("cpu"))
        out, batch_logs = [], []
        cb = mmu.src.callbacks.LambdaCallback(on_epoch_end=lambda e, l: out.append(dict(l)),
                                              on_batch_end=lambda b, l: batch_logs.append(dict(l)))
        trainer.train_loop(train, valid_generator=val, epochs=2, steps_per_epoch=len(train),
                           validation_steps=len(val), callbacks=[cb], scheduler_step_on="batch",
                           metrics_every=every)
        logs[every] = out
        assert len(out) == 2 and len(batch_logs) == 2 * len(train)
        assert ("deferred" in batch_logs[0]) == (every > 1)
    for every in (3, 100):
        for a, b in zip(logs[1], logs[every]):
            assert a["loss"] == pytest.approx(b["loss"], rel=1e-6) and a["acc"] == pytest.approx(b["acc"], abs=1e-4)
            assert a["val_loss"] == pytest.approx(b["val_loss"], rel=1e-6)
    # NaN loss stops training at the end of the epoch in both protocols
    for every in (1, 4):
        torch.manual_seed(0)
        train, _, _ = mmu.dataset.get_synthetic_flava(4, 12, 8, 8, l_img=3, l_txt=2, dim=3, num_classes=3)
        net = Tiny()
        net.poison = True
        trainer = mmu.Model_(net, torch.optim.SGD(net.parameters(), lr=0.1), None,
                             partial(mmu.dataset.data_forming_func_transformer, model_type="MultiHead"),
                             metrics=[acc], verbose=False)
        trainer.to(torch.device("cpu"))
        seen = []
        trainer.train_loop(train, epochs=5, steps_per_epoch=len(train), scheduler_step_on="batch",
                           callbacks=[mmu.src.callbacks.LambdaCallback(on_epoch_end=lambda e, l: seen.append(e))],
                           metrics_every=every)
        assert seen == [1]


def test_trainer_vilt_branch_with_transformers_vilt(mmu):
    """The ``vilt`` branch of Model_ (reference src/framework.py:163-169, 263-304; set up by
    train.py:164-182) driving the SAME class the reference wraps --
    transformers.ViltForImagesAndTextClassification, here with a tiny random-init config (the
    "dandelin/vilt-b32-mlm" checkpoint is unavailable offline): dict batches, ``model(**batch)``,
    ``outputs.loss`` / ``.logits``, labels from the batch, (B, C) metrics, gradient accumulation,
    epoch-wise ReduceLROnPlateau on val_acc."""
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.ViltConfig(hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64,
                                  image_size=32, patch_size=16, num_labels=3, num_images=1, vocab_size=100,
                                  max_position_embeddings=16)
    model = transformers.ViltForImagesAndTextClassification(cfg)
    g = torch.Generator().manual_seed(1)

    def batches(n):
        return [dict(input_ids=torch.randint(0, 100, (4, 8), generator=g),
                     attention_mask=torch.ones(4, 8, dtype=torch.long),
                     token_type_ids=torch.zeros(4, 8, dtype=torch.long),
                     pixel_values=torch.randn(4, 1, 3, 32, 32, generator=g),
                     labels=torch.randint(0, 3, (4,), generator=g)) for _ in range(n)]

    def acc(y_pred, y_true, eval, dummy_dim=False):         # train.py:119-130 with dummy_dim=False
        assert not dummy_dim and y_pred.dim() == 2
        return (y_pred.argmax(1) == y_true).float().mean() * 100
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "max", patience=1, factor=0.5)
    steps = []
    orig = opt.step
    opt.step = lambda *a, **k: (steps.append(1), orig(*a, **k))[1]
    trainer = mmu.Model_(model, opt, sched, None, metrics=[acc], verbose=False)
    trainer.to(torch.device("cpu"))
    logs = []
    trainer.train_loop(batches(4), valid_generator=batches(2), epochs=2, steps_per_epoch=4, validation_steps=2,
                       callbacks=[mmu.src.callbacks.LambdaCallback(on_epoch_end=lambda e, l: logs.append(dict(l)))],
                       scheduler_step_on="epoch", scheduler_metric="val_acc", vilt=True,
                       gradient_accumulation_steps=2, auc=False)
    assert len(logs) == 2 and len(steps) == 4                # 8 batches, one optimizer step per 2
    for k in ("loss", "acc", "val_loss", "val_acc"):
        assert k in logs[0] and np.isfinite(logs[1][k]), k
    out = trainer.eval_loop(batches(2), "test", vilt=True)
    assert set(out) >= {"test_loss", "test_acc"}


def test_trainer_mmbt_branch_with_a_plain_torch_model(mmu, monkeypatch):
    """The ``mmbt`` branches of Model_ (reference src/framework.py:246-304, :172-176): ``model(*x)``,
    per-epoch freeze flags on ``enc.img_encoder`` / ``enc.encoder``, gradient-accumulation stepping,
    (B, C) logits with ``dummy_dim=False`` metrics, epoch-wise scheduler on ``val_acc``."""
    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.img_encoder = torch.nn.Linear(4, 4)
            self.encoder = torch.nn.Linear(8, 8)
    class TinyMM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = Enc()
            self.emb = torch.nn.Embedding(20, 4)
            self.clf = torch.nn.Linear(8, 2)
        def forward(self, txt, mask, segment, img):
            t = (self.emb(txt) * mask[..., None]).sum(1)
            return self.clf(self.enc.encoder(torch.cat([t, self.enc.img_encoder(img)], -1)))
        def compute_loss(self, y_hat, y, eval=False):
            return torch.nn.functional.cross_entropy(y_hat, y)
    def acc(y_pred, y_true, eval, dummy_dim=False):
        assert dummy_dim is False and y_pred.dim() == 2
        return (y_pred.argmax(1) == y_true).float().mean() * 100
    torch.manual_seed(0)
    def batches(n):
        return [((torch.randint(0, 20, (4, 5)), torch.ones(4, 5, dtype=torch.long),
                  torch.ones(4, 5, dtype=torch.long), torch.randn(4, 4)), torch.randint(0, 2, (4,)))
                for _ in range(n)]
    net = TinyMM()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "max", patience=1, factor=0.5)
    seen = []
    step = opt.step
    opt.step = lambda *a, **k: (seen.append((net.enc.img_encoder.weight.requires_grad,
                                             net.enc.encoder.weight.requires_grad)), step(*a, **k))[1]
    trainer = mmu.Model_(net, opt, sched, lambda x, y, phase="train": (x, y), metrics=[acc], verbose=False)
    trainer.to(torch.device("cpu"))
    # the product's AUROC counts pairs on the GPU and refuses to run without one; this host-logic
    # test stands the oracle in for it
    from oracle import rank_stats
    if not torch.cuda.is_available():
        with pytest.raises(mmu._lib.MMUError):
            mmu.src.metrics.auroc(torch.tensor([0, 1]), torch.tensor([0.2, 0.7]))
    monkeypatch.setattr(mmu.src.metrics, "auroc",
                        lambda lab, sc: rank_stats.auroc(lab.numpy(), sc.numpy()))
    H = {}
    cbs = [mmu.src.callbacks.LambdaCallback(on_epoch_end=lambda e, l: H.setdefault("logs", []).append(dict(l)))]
    trainer.train_loop(batches(4), valid_generator=batches(2), epochs=2, steps_per_epoch=4, validation_steps=2,
                       callbacks=cbs, scheduler_step_on="epoch", scheduler_metric="val_acc", mmbt=True,
                       freeze_img=2, freeze_txt=0, gradient_accumulation_steps=2, auc=True)
    # 4 batches / accumulation 2 -> 2 optimizer steps per epoch; image encoder frozen in epoch 1 only
    assert seen == [(False, True)] * 2 + [(True, True)] * 2
    assert len(H["logs"]) == 2 and {"loss", "acc", "val_loss", "val_acc", "val_auc"} <= set(H["logs"][0])


def _gloo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import mmu_b200
    par, lib = mmu_b200.parallel, mmu_b200._lib
    # metric accumulators: integer bins add exactly, doubles add
    acc = torch.zeros(lib.ACC_WORDS, dtype=torch.int64)
    acc[lib.ACC_OFF["conf_count"] + 3] = 10 + rank
    acc[lib.ACC_OFF["n_samples"]] = 100 * (rank + 1)
    acc[lib.ACC_INT_WORDS:].view(torch.float64)[0] = 0.5 + rank
    par.all_reduce_accum(acc)
    assert int(acc[lib.ACC_OFF["conf_count"] + 3]) == 21 and int(acc[lib.ACC_OFF["n_samples"]]) == 300
    assert float(acc[lib.ACC_INT_WORDS:].view(torch.float64)[0]) == 2.0
    # bucketed gradient all-reduce over stage ranges + folded 1/world averaging
    flat = torch.arange(40, dtype=torch.float32) * (rank + 1)
    works = par.all_reduce_ranges(flat, [(30, 40), (10, 30), (0, 10)], async_op=True)
    for w in works:
        w.wait()
    assert torch.equal(flat * (1.0 / world), torch.arange(40, dtype=torch.float32) * 1.5)
    p = torch.full((8,), float(rank))
    par.broadcast_flat(p, 0)
    assert float(p.abs().sum()) == 0.0
    # post-hoc scoring accumulator: Pearson sums (doubles) and per-variant counts (uint64) merge
    ph = mmu_b200.metrics.PosthocMeter("cpu", n_repeats=2)
    words = ph.accum.view(torch.int64)
    words[:10].view(torch.float64)[:] = torch.arange(10, dtype=torch.float64) + rank
    words[10] = 50 * (rank + 1)          # n_samples
    words[11 + 3] = 7 + rank             # correct[3]
    ph.all_reduce()
    assert torch.equal(words[:10].view(torch.float64), 2 * torch.arange(10, dtype=torch.float64) + 1)
    assert int(words[10]) == 150 and int(words[11 + 3]) == 15
    assert ph.compute()["n_samples"] == 150
    # rank statistics need every sample's score: uneven shards are gathered in rank order
    mine = torch.arange((3 + 2 * rank) * 4, dtype=torch.float32).reshape(3 + 2 * rank, 4) + 100 * rank
    full = par.all_gather_rows(mine)
    want = torch.cat([torch.arange((3 + 2 * r) * 4, dtype=torch.float32).reshape(3 + 2 * r, 4) + 100 * r
                      for r in range(world)])
    assert torch.equal(full, want)
    lab = par.all_gather_rows(torch.full((2 + rank,), rank, dtype=torch.int64))
    assert lab.tolist() == [0, 0, 1, 1, 1]
    # flat-buffer owners of the MMBT path (trunk + image encoder): broadcast + gradient sum
    class Owner:
        def __init__(self, n):
            self._flat = torch.full((n,), float(rank + 1))
            self._flat_grad = torch.arange(n, dtype=torch.float32) * (rank + 1)
            self._stats = torch.full((3,), float(rank))
    class Opt:
        _owners = [Owner(5), Owner(7)]
    sync = par.FlatGradSync.attach(Opt)
    assert Opt.grad_scale == 0.5 and all(float(o._flat.sum()) == o._flat.numel() for o in Opt._owners)
    assert all(float(o._stats.abs().sum()) == 0.0 for o in Opt._owners)
    sync.all_reduce_grads()
    assert torch.equal(Opt._owners[1]._flat_grad, torch.arange(7, dtype=torch.float32) * 3)
    # sample sharding: contiguous, disjoint, covering
    b, e = par.shard_range(1000003, rank, world)
    t = torch.tensor([b, e])
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert int(gathered[0][0]) == 0 and int(gathered[-1][1]) == 1000003
    assert all(int(gathered[i][1]) == int(gathered[i + 1][0]) for i in range(world - 1))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_packed_store_round_trip(mmu, tmp_path):
    """pack_flava_encodings / PackedFlavaDataset serve exactly the items the per-sample files
    of the reference would (src/dataset.py:206-213), ragged text lengths included."""
    src = mmu.dataset.SyntheticFlavaDataset(9, l_img=5, l_txt=7, dim=8, num_classes=4, seed=3, ragged=True)
    items = [src[i] for i in range(len(src))]
    mmu.dataset.pack_flava_encodings(((a, b, int(c)) for a, b, c in items), str(tmp_path / "store"))
    ds = mmu.dataset.PackedFlavaDataset(str(tmp_path / "store"))
    assert len(ds) == len(items)
    for i, (a, b, c) in enumerate(items):
        ia, ib, ic = ds[i]
        assert torch.equal(ia, a) and torch.equal(ib, b) and torch.equal(ic, c)
    # host collate of the packed items == host collate of the originals (reference behaviour)
    (pi, pt), py = mmu.dataset.collate_fn_flava([ds[i] for i in (0, 3, 4)])
    (oi, ot), oy = mmu.dataset.collate_fn_flava([items[i] for i in (0, 3, 4)])
    assert torch.equal(pi, oi) and torch.equal(pt, ot) and torch.equal(py, oy)


def test_mmbt_state_dict_keys_order_and_strict_load(mmu, golden):
    """MultimodalBertClf: the reference's state_dict keys (incl. the embedding tensors
    ImageBertEmbeddings shares with the text side, src/mmbt.py:51-55), named_parameters() order and
    a strict load of a reference checkpoint; host-side index lists of the forward variants."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_mmbt import make_args
    c = golden("mmbt_small.pt")["fp32_small"]
    m = mmu.MultimodalBertClf(make_args(c["cfg"], "fp32"))
    assert sorted(m.state_dict()) == sorted(c["state_dict_keys_all"])
    assert [k for k, _ in m.named_parameters()] == c["named_parameters"]
    m.load_state_dict(c["state_dict"], strict=True)
    assert m.enc.img_embeddings.word_embeddings.weight is m.enc.txt_embeddings.word_embeddings.weight
    w = m.enc.encoder.layer[0].attention.self.query.weight
    assert torch.equal(w, c["state_dict"]["enc.encoder.layer.0.attention.self.query.weight"])
    for modal, d in c["control"].items():
        torch.manual_seed(d["seed"])
        num = c["cfg"]["n_img"] + 1 if modal == "image" else c["cfg"]["S_txt"]
        ind = m.control_indices(c["cfg"]["S_txt"] + c["cfg"]["n_img"] + 2, num)
        assert ind == [int(i) for i in d["indices"]]  # bit-exact with the reference's draw
    with pytest.raises(mmu._lib.MMUError):  # no CPU path
        m(c["txt"], c["mask"], c["segment"], c["img_tokens"])


def test_graphed_train_step_host_logic(mmu):
    """graphs.GraphedTrainStep without a GPU: what it refuses to capture and what re-captures.
    (The capture itself is exercised by tests/test_gpu_model.py::test_cuda_graph_train_step_equals_eager.)"""
    import types
    G = mmu.graphs.GraphedTrainStep

    class WithFB(torch.nn.Linear):
        def forward_backward(self, x, y):
            raise AssertionError("not reached")

    net = WithFB(3, 2)
    sgd = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    tr = types.SimpleNamespace(model=net, optimizer=sgd, metrics=[], device=torch.device("cpu"))
    g = G(tr)
    x, y = torch.zeros(4, 3), torch.zeros(4, dtype=torch.long)
    k0 = g._key([x], y)
    assert g._key([x], y) == k0
    assert g._key([torch.zeros(5, 3)], torch.zeros(5, dtype=torch.long)) != k0      # batch shape
    assert g._key([x.double()], y) != k0                                              # dtype
    sgd.param_groups[0]["lr"] = 0.05                                                  # ReduceLROnPlateau
    assert g._key([x], y) != k0
    assert g._key([x, None], y)[0][1] is None                                         # absent modality
    # a model without the autograd-free entry point, optimisers with host-side step scalars
    with pytest.raises(TypeError):
        G(types.SimpleNamespace(model=torch.nn.Linear(3, 2), optimizer=sgd, metrics=[], device=None))
    FusedAdamW = type("FusedAdamW", (), {"param_groups": []})
    with pytest.raises(ValueError):
        G(types.SimpleNamespace(model=net, optimizer=FusedAdamW(), metrics=[], device=None))
    adam = torch.optim.Adam(net.parameters(), lr=1e-3)           # capturable=False by default
    with pytest.raises(ValueError):
        G(types.SimpleNamespace(model=net, optimizer=adam, metrics=[], device=None))
    G(types.SimpleNamespace(model=net, optimizer=torch.optim.Adam(net.parameters(), capturable=True),
                            metrics=[], device=None))


# ---- dropout masks: the product's mask function compiled for the HOST against the oracle ---------

_DROPOUT_SHIM = r"""
#define __host__
#define __device__
#define __forceinline__ inline
#include "dropout.cuh"
extern "C" void site_params(float p, unsigned long long seed, unsigned int site, unsigned int* out3, float* scale) {
  const mmu::dropout::Site s = mmu::dropout::make_site(p, seed, site);
  out3[0] = s.lo; out3[1] = s.hi; out3[2] = s.thresh; *scale = s.scale;
}
extern "C" void keep_range(float p, unsigned long long seed, unsigned int site, unsigned int first,
                           unsigned int n, unsigned char* keep, float* mult) {
  const mmu::dropout::Site s = mmu::dropout::make_site(p, seed, site);
  for (unsigned int i = 0; i < n; ++i) { keep[i] = s.keep(first + i); mult[i] = s.mult(first + i); }
}
"""


@pytest.fixture(scope="module")
def dropout_host(tmp_path_factory):
    """csrc/dropout.cuh is __host__ __device__ code: built here with g++ as the product's own mask
    function (no GPU, no oracle involved in producing it)."""
    import ctypes
    import subprocess
    d = tmp_path_factory.mktemp("dropout_host")
    src = d / "shim.cpp"
    src.write_text(_DROPOUT_SHIM)
    so = d / "libdropout_host.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17",
                           "-I", os.path.join(ROOT, "multi-modal-uncertainty_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = ctypes.CDLL(str(so))
    lib.site_params.argtypes = [ctypes.c_float, ctypes.c_ulonglong, ctypes.c_uint,
                                ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_float)]
    lib.keep_range.argtypes = [ctypes.c_float, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint,
                               ctypes.c_void_p, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("p", [0.0, 0.1, 0.25, 0.5, 0.9, 1.0])
def test_dropout_mask_function_of_the_kernels_equals_the_oracle_restatement(dropout_host, p):
    """Bit-exact: site seed (splitmix64), threshold, hash and multiplier of csrc/dropout.cuh against
    oracle/dropout.py -- over several seeds / sites, a counter range that wraps 2^32, and the
    extreme probabilities (p = 0: everything kept, scale 1; p = 1: threshold saturates)."""
    import ctypes
    from oracle import dropout
    n = 1 << 16
    for seed, site, first in ((0, 0, 0), (42, 3, 12345), (2 ** 63 + 11, 70, 2 ** 32 - 1000), (2 ** 64 - 1, 2 ** 32 - 1, 7)):
        out3, scale = (ctypes.c_uint * 3)(), ctypes.c_float()
        dropout_host.site_params(p, seed, site, out3, ctypes.byref(scale))
        ss = dropout.site_seed(seed, site)
        assert (out3[0], out3[1]) == (ss & 0xFFFFFFFF, ss >> 32)
        assert out3[2] == dropout.threshold(p)
        keep = np.zeros(n, dtype=np.uint8)
        mult = np.zeros(n, dtype=np.float32)
        dropout_host.keep_range(p, seed, site, first, n, keep.ctypes.data, mult.ctypes.data)
        ref_keep = dropout.keep_mask(p, seed, site, n, offset=first)
        assert torch.equal(torch.from_numpy(keep).bool(), ref_keep)
        if p < 1.0:
            assert torch.equal(torch.from_numpy(mult), dropout.multiplier(p, seed, site, n, offset=first))
        if p == 0.0:
            assert keep.all() and scale.value == 1.0


def test_dropout_masks_have_the_statistics_of_nn_dropout():
    """Statistical parity with nn.Dropout's definition (src/model.py:195-201): each element kept
    independently with probability 1 - p, survivors scaled by 1 / (1 - p) so that the expectation is
    the identity; different sites / seeds give independent masks; the mask is a pure function of
    (seed, site, index)."""
    from oracle import dropout
    n = 1 << 20
    for p in (0.1, 0.25, 0.5):
        k = dropout.keep_mask(p, 1234, 5, n).double()
        sigma = (p * (1 - p) / n) ** 0.5
        assert abs(float(k.mean()) - (1 - p)) < 4 * sigma
        m = dropout.multiplier(p, 1234, 5, n)
        assert abs(float(m.double().mean()) - 1.0) < 4 * sigma / (1 - p)
        assert set(m.unique().tolist()) == {0.0, float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))}
        c = k - k.mean()
        var = float((c * c).mean())
        for lag in (1, 2, 7, 768, 3072):   # neighbours along a row and down a column of the GEMM outputs
            assert abs(float((c[:-lag] * c[lag:]).mean()) / var) < 5 / n ** 0.5
        for other in (dropout.keep_mask(p, 1234, 6, n), dropout.keep_mask(p, 1235, 5, n)):   # site, seed
            o = other.double() - other.double().mean()
            assert abs(float((c * o).mean()) / var) < 5 / n ** 0.5
        assert torch.equal(dropout.keep_mask(p, 1234, 5, n), dropout.keep_mask(p, 1234, 5, n))
        assert torch.equal(dropout.keep_mask(p, 1234, 5, 1000, offset=500), dropout.keep_mask(p, 1234, 5, n)[500:1500])


def test_c_abi_argument_errors_need_no_device(mmu):
    """Error behaviour of the boundary (include/mmu_b200.h: 'return value: 0 on success, a negative
    MMU_ERR_* code otherwise; nothing throws'): every code has its own message, and null configs /
    null tensors are rejected with MMU_ERR_ARG by the host-side validation, before any CUDA call --
    so the check runs on a machine without a GPU."""
    import ctypes as C
    L = mmu._lib.lib
    hdr = open(os.path.join(ROOT, "include", "mmu_b200.h")).read()
    codes = {n: int(v) for n, v in re.findall(r"#define (MMU_(?:OK|ERR_[A-Z]+)) \(?(-?\d+)\)?", hdr)}
    assert codes["MMU_OK"] == 0 and len(codes) == 8 and sorted(codes.values()) == list(range(-7, 1))
    msgs = {c: L.mmu_error_string(c).decode() for c in codes.values()}
    assert all(msgs.values()) and len(set(msgs.values())) == len(msgs)
    assert L.mmu_error_string(-99).decode() not in set(msgs.values())
    ARG = codes["MMU_ERR_ARG"]
    for count in (L.mmu_flava_param_count, L.mmu_resnet_param_count, L.mmu_mmbt_param_count, L.mmu_imgenc_param_count):
        assert count(None) == ARG
    assert L.mmu_mask_gather_tokens(None, None, 1, 1, 1, 1, None, 0, None, 0, 0, None) == ARG
    assert L.mmu_layernorm_fwd(None, None, None, None, 0, None, None, 4, 4, None) == ARG
    assert L.mmu_cast_f32_to_bf16(None, None, 16, None) == ARG
    assert L.mmu_ragged_pad(None, None, None, 1, 1, 1, None) == ARG
    assert L.mmu_heads_uncertainty_epilogue(None, None, 1, 1, 1, 1, 1, 1, 0.0, None, None, None, None, None) == ARG
    assert L.mmu_adamw_flat_step(None, None, None, None, None, 16, 0.1, 0.9, 0.98, 1e-9, 0.0, 1, 1.0, None) == ARG
    assert L.mmu_flava_forward(None, None, None, None, 0, 0, None, None) == ARG
    assert L.mmu_gemm(0, None, 0, 0, None, 0, 0, 4, 4, 4, 1, None, None) == ARG
    e = mmu._lib.GemmEpilogue()
    e.drop_p = 1.5   # nn.Dropout accepts p in [0, 1]; the epilogue needs 1 / (1 - p) finite
    assert L.mmu_gemm(0, None, 0, 0, None, 0, 0, 4, 4, 4, 1, C.byref(e), None) == ARG
    with pytest.raises(mmu._lib.MMUError, match="bad argument"):
        mmu._lib.check(ARG, "probe")


def test_fmnist_view_format_sweep_and_model_table(mmu, golden):
    """Either side of the FashionMNIST path, against the reference run in the build container
    (tests/golden/fmnist_views.pt): ``dataset.quarter_views`` == QuarterCrop + ToTensor (batched and
    per sample, bit-exact); ``robustness.run_view_robustness`` == the reference script's own sweep
    statements for a multi-head model (view zero-filled) and for ``single-model-weight-sharing``
    (view removed, labels repeated as the reference saves them) -- host shaping only, with the
    golden's linear stand-in model and without the device-side meters; and the names the
    reference's FashionMNIST scripts import from ``src.model`` resolve."""
    import importlib
    c = golden("fmnist_views.pt")
    imgs = c["images_u8"].float().div(255).unsqueeze(1)
    assert torch.equal(mmu.dataset.quarter_views(imgs), c["quarters"])
    assert torch.equal(mmu.dataset.quarter_views(imgs[2]), c["quarters"][2])
    with pytest.raises(ValueError):
        mmu.dataset.quarter_views(torch.zeros(1, 27, 28))
    C_ = 10

    class Lin(torch.nn.Module):
        def __init__(self, W, heads):
            super().__init__()
            self.W, self.heads = W, heads

        def forward(self, x):
            return (x.reshape(x.shape[0], -1) @ self.W).view(-1, self.heads, C_)

    for mt, model in (("MultiHead", Lin(c["W4"], 4)), ("single-model-weight-sharing", Lin(c["W1"], 1))):
        P, labels, per_view = mmu.robustness.run_view_robustness(model, c["valid"], "cpu", model_type=mt, metrics=False)
        assert per_view == []
        assert torch.equal(torch.from_numpy(P), c[mt]["outputs"])
        assert torch.equal(torch.from_numpy(labels), c[mt]["labels"])
        # the plain prediction dump (eval_prediction_saving.py:77-104) with the same stand-in model
        P, labels, summary = mmu.robustness.run_predictions(model, c["valid"], "cpu", model_type=mt, metrics=False)
        assert summary is None
        assert torch.equal(torch.from_numpy(P), c["predictions/" + mt]["outputs"])
        assert torch.equal(torch.from_numpy(labels), c["predictions/" + mt]["labels"])
    x = c["valid"][0][0]
    assert mmu.robustness.view_sweep_inputs(x, 1, "single-model-weight-sharing").shape == (x.shape[0] * 3, 1, 14, 14)
    # `from src.model import MIMOResNet, model_configure, MIMOTransfomer` (train_fashionmnist.py:17)
    src_model = importlib.import_module("multi-modal-uncertainty_b200.src.model")
    assert src_model.MIMOResNet is mmu.MIMOResNet and src_model.MIMOTransfomer is mmu.MIMOTransfomer
    assert src_model.model_configure == {"Vanilla": (4, 1), "MIMO-shuffle-instance": (4, 4), "MIMO-shuffle-view": (4, 4),
                                         "MultiHead": (4, 4), "MIMO-shuffle-all": (4, 4),
                                         "single-model-weight-sharing": (1, 1)}
    assert mmu.model_configure is src_model.model_configure
    with pytest.raises(AttributeError):
        src_model.no_such_name


def test_resume_from_last_epoch_checkpoint_as_train_py_does(mmu, tmp_path):
    """The reference's resume path (train.py:273-283): reload ``model_last_epoch.pt`` strictly into a
    freshly constructed model, rebuild ``H`` from ``history.csv``, continue at
    ``epoch_start = len(H['epoch']) + 1`` with the default callbacks -- the history keeps growing in
    the same file, the per-epoch checkpoints continue their numbering, and the best-validation
    checkpoint is only replaced by a better epoch."""
    import pandas as pd
    from functools import partial

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(6, 3 * 2)

        def forward(self, x):
            img, txt = x
            return self.fc(torch.cat([img.mean(1), txt.mean(1)], -1)).view(-1, 2, 3)

        def compute_loss(self, y_hat, y, eval=False):
            y_hat = y_hat.mean(1) if eval else y_hat.reshape(-1, 3)
            return torch.nn.functional.cross_entropy(y_hat, y.reshape(-1))

    def acc(y_pred, y_true, eval, dummy_dim=False):
        y_pred = y_pred.mean(1) if eval else y_pred.reshape(-1, 3)
        return (y_pred.argmax(1) == y_true.reshape(-1)).float().mean() * 100

    tl = mmu.src.training_loop
    train, val, test = mmu.dataset.get_synthetic_flava(4, 12, 8, 8, l_img=3, l_txt=2, dim=3, num_classes=3)

    def run(net, H, epoch_start, epochs):
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
        cbs = tl._construct_default_callbacks(net, opt, H, str(tmp_path), checkpoint_monitor="val_acc")
        for cb in cbs:
            cb.set_save_path(str(tmp_path)); cb.set_model(net, ignore=False); cb.set_optimizer(opt)
        trainer = mmu.Model_(net, opt, sched, partial(mmu.dataset.data_forming_func_transformer, model_type="MultiHead"),
                             metrics=[acc], verbose=False)
        trainer.to(torch.device("cpu"))
        trainer.train_loop(train, valid_generator=val, test_generator=test, epochs=epochs, epoch_start=epoch_start,
                           steps_per_epoch=len(train), validation_steps=len(val), test_steps=len(test),
                           callbacks=cbs, scheduler_step_on="batch", scheduler_metric=None)

    torch.manual_seed(0)
    first = Tiny()
    run(first, {}, 1, 2)
    saved = {k: v.clone() for k, v in first.state_dict().items()}

    # ---- what train.py does under --resume
    torch.manual_seed(1)
    net = Tiny()
    checkpoint = torch.load(tmp_path / "model_last_epoch.pt")
    assert set(checkpoint) == {"model", "optimizer"}
    net.load_state_dict(checkpoint["model"])
    assert all(torch.equal(v, saved[k]) for k, v in net.state_dict().items())
    H = pd.read_csv(tmp_path / "history.csv")
    H = {col: list(H[col].values) for col in H.columns if col != "Unnamed: 0"}
    epoch_start = len(H["epoch"]) + 1
    assert epoch_start == 3
    best_before = torch.load(tmp_path / "model_best_val.pt")["model"]
    run(net, H, epoch_start, 4)

    H2 = pd.read_csv(tmp_path / "history.csv")
    assert list(H2["epoch"]) == [1, 2, 3, 4] and len(H2["val_acc"]) == 4 and not H2["loss"].isna().any()
    for e in (1, 2, 3, 4):
        assert os.path.exists(tmp_path / f"model_epoch_{e}.pt")
    last = torch.load(tmp_path / "model_last_epoch.pt")["model"]
    assert all(torch.equal(last[k], v) for k, v in net.state_dict().items())
    # a ModelCheckpoint created for the resumed run starts from its own best (= -inf), as in the
    # reference: the file is rewritten by the first resumed epoch and afterwards only by a better one
    best_after = torch.load(tmp_path / "model_best_val.pt")["model"]
    resumed_best_epoch = 3 if H2["val_acc"][3] <= H2["val_acc"][2] else 4
    expect = torch.load(tmp_path / f"model_epoch_{resumed_best_epoch}.pt")["model"]
    assert all(torch.equal(best_after[k], expect[k]) for k in expect)
    assert set(best_before) == set(best_after)


def test_callback_accessors_and_checkpoint_pickling(mmu, tmp_path):
    """Reference src/callbacks.py:82-125 (setter / getter pairs of ``Callback``) and :217-226
    (``ModelCheckpoint`` pickles without its model / optimizer)."""
    import pickle
    cbm = mmu.src.callbacks
    cb = cbm.Callback()
    for name, value in (("meta_data", {"a": 1}), ("save_path", "p"), ("optimizer", object()), ("params", {"epochs": 3}),
                        ("dataloader", [1, 2])):
        getattr(cb, "set_" + name)(value)
        assert getattr(cb, "get_" + name)() is value
    lin = torch.nn.Linear(2, 2)
    cb.set_model(lin)                      # ignore=True is the default: the reference's "trick"
    assert not hasattr(cb, "model")
    cb.set_model(lin, ignore=False)
    assert cb.get_model() is lin
    ck = cbm.ModelCheckpoint(str(tmp_path / "best.pt"), monitor="val_acc", save_best_only=True, mode="max")
    ck.set_model(lin, ignore=False)
    ck.set_optimizer(torch.optim.SGD(lin.parameters(), lr=0.1))
    ck.on_epoch_end(1, {"val_acc": 50.0})
    blob = pickle.dumps(ck)
    assert b"Linear" not in blob
    ck2 = pickle.loads(blob)
    assert ck2.best == 50.0 and ck2.monitor == "val_acc" and not hasattr(ck2, "model")
    ck.on_epoch_end(2, {"val_acc": 40.0})  # not better: file untouched
    t = os.path.getmtime(tmp_path / "best.pt")
    ck.on_epoch_end(3, {"val_acc": 60.0})
    assert ck.best == 60.0 and os.path.getmtime(tmp_path / "best.pt") >= t


def test_bench_accounting_and_reference_arm_contract(monkeypatch, capsys):
    """bench.py on the CPU: (1) the algorithmic FLOP count `roofline` / `whole_step` divide by is
    SURVEY 8d's formula at BASELINE.json's configuration (10.63 GFLOP forward, ~31.9 GFLOP train
    per sample, all positions as written); the sweep's token counts follow the 10-level grid;
    (2) the `--impl reference` line keeps the contract (same metric / unit / config / direction as
    the product arm, `impl`, zero-copy `e2e`, a `cpu_baseline` of this run, bounded steps) and only
    rank 0 prints it; (3) without a GPU the product arm refuses to run instead of falling back."""
    import argparse
    import json
    import bench
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.CFG["B"] == 128 and (bench.CFG["l_img"], bench.CFG["l_txt"]) == (197, 40)
    assert (bench.CFG["D"], bench.CFG["layers"], bench.CFG["E"], bench.CFG["C"]) == (768, 3, 5, 101)
    assert "samples/sec" in cfg["metric"] or "samples" in json.dumps(cfg)
    f = bench.fwd_flops_per_sample(197, 40, 128)
    L, D = 237, 768
    assert f == 2 * L * D * D + 3 * (24 * L * D * D + 4 * L * 128 * D) + 2 * 5 * D * 101
    assert abs(f / 1e9 - 10.63) < 0.01
    total, train, sweep = bench.step_flops(128)
    assert total == train + sweep and abs(train / 128 / 1e9 - 31.9) < 0.4
    counts = bench.level_token_counts()
    assert len(counts) == 10 and counts[0] == (0, 40) and counts[-1] == (197, 0)
    assert all(0 <= a <= 197 and 0 <= b <= 40 for a, b in counts)
    assert [a for a, _ in counts] == sorted(a for a, _ in counts)
    assert sweep == 128 * sum(bench.fwd_flops_per_sample(a, b, 128) for a, b in counts)

    calls = []

    def fake_cpu_reference(steps, warmup, B=None):
        calls.append((steps, warmup))
        return {"value": 8.5, "unit": "samples/s", "cores": 16, "kind": "port", "sample": "stub", "s_per_step": 15.0}

    monkeypatch.setattr(bench, "cpu_reference", fake_cpu_reference)
    args = argparse.Namespace(steps=20, warmup=3, gpus=1)
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert calls == [(3, 1)]                                   # bounded: the whole run ends within minutes
    assert line["impl"] == "reference" and line["metric"] == "train+robustness-eval samples/sec"
    assert line["unit"] == "samples/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] == 8.5 and line["e2e"] == {"value": 8.5, "unit": "samples/s", "h2d_bytes_per_step": 0,
                                                     "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 16
    assert line["config"]["workload"] == bench.WORKLOAD_NAME and line["config"]["per_gpu_batch"] == 128
    monkeypatch.setenv("RANK", "1")                            # other ranks exit without work or output
    bench.run_reference(args)
    assert capsys.readouterr().out == "" and calls == [(3, 1)]
    if not torch.cuda.is_available():
        monkeypatch.setattr(sys, "argv", ["bench.py"])
        with pytest.raises(SystemExit, match="no CPU fallback"):
            bench.main()


def test_product_mmbt_collate_bit_exact(mmu, golden):
    """dataset.collate_fn (the MMBT batch format) against the reference's own collate_fn, dtypes
    included; segment == mask on every real token, which is why the reference's ``model(*x)``
    argument order (segment in the mask slot) is harmless."""
    c = golden("mmbt_collate.pt")
    (txt, segment, mask, img), tgt = mmu.dataset.collate_fn(c["rows"])
    for got, key in ((txt, "txt"), (segment, "segment"), (mask, "mask"), (img, "img"), (tgt, "tgt")):
        assert got.dtype == c[key].dtype and torch.equal(got, c[key]), key
    assert torch.equal(segment, mask)


def test_epoch_stepped_plateau_scheduler_as_train_fashionmnist_does(mmu):
    """train_fashionmnist.py:111-130,196-213: SGD + ``ReduceLROnPlateau`` stepped ONCE PER EPOCH on
    ``epoch_log[scheduler_metric]`` (src/framework.py:337-338), four-view batches through
    ``data_forming_func(model_type='MultiHead')``, ``acc`` with ``dummy_dim=True``."""
    from functools import partial

    class TinyViews(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(4 * 14 * 14, 4 * 10)

        def forward(self, x):                      # (B, 4, 1, 14, 14) -> (B, 4, 10)
            return self.fc(x.reshape(x.shape[0], -1)).view(-1, 4, 10)

        def compute_loss(self, y_hat, y, eval=False):
            y_hat = y_hat.mean(1) if eval else y_hat.reshape(-1, 10)
            return torch.nn.functional.cross_entropy(y_hat, y.reshape(-1))

    def acc(y_pred, y_true, eval, dummy_dim=False):     # train.py:119-130 on the host (mmu.acc runs on the device)
        if dummy_dim:
            y_pred, y_true = (y_pred.mean(1), y_true) if eval else (y_pred.reshape(-1, y_pred.shape[2]), y_true.reshape(-1))
        return (y_pred.max(1)[1] == y_true).float().mean() * 100

    g = torch.Generator().manual_seed(0)
    batches = [(torch.rand(6, 4, 1, 14, 14, generator=g), torch.randint(0, 10, (6,), generator=g)) for _ in range(3)]
    net = TinyViews()
    opt = torch.optim.SGD(net.parameters(), lr=0.1, weight_decay=0.001, momentum=0.9)
    seen = []

    class Plateau(torch.optim.lr_scheduler.ReduceLROnPlateau):
        def step(self, metrics, *a, **k):
            seen.append(float(metrics))
            return super().step(metrics, *a, **k)

    sched = Plateau(opt, mode="min", factor=0.1, patience=0, threshold=1e9, threshold_mode="abs")  # never "better"
    logs = []
    trainer = mmu.Model_(net, opt, sched, partial(mmu.dataset.data_forming_func, model_type="MultiHead"),
                         metrics=[acc], verbose=False)
    trainer.to(torch.device("cpu"))
    cb = mmu.src.callbacks.LambdaCallback(on_epoch_end=lambda e, l: logs.append(dict(l)))
    trainer.train_loop(batches, valid_generator=batches[:2], test_generator=batches[:1], epochs=3,
                       steps_per_epoch=3, validation_steps=2, test_steps=1, callbacks=[cb],
                       scheduler_step_on="epoch", scheduler_metric="val_loss")
    assert len(seen) == 3 and seen == [l["val_loss"] for l in logs]          # one step per epoch, on val_loss
    # epoch 1 sets the best value; epochs 2 and 3 are "not better" with patience 0: two reductions
    assert abs(opt.param_groups[0]["lr"] - 0.1 * 0.1 ** 2) < 1e-12
    assert {"epoch", "loss", "acc", "val_loss", "val_acc", "test_loss", "test_acc"} <= set(logs[0])


def test_bertadam_accepts_the_fmnist_transformer_configuration(mmu):
    """train_fashionmnist.py:91-109: ``MIMOTransfomer`` + ``BertAdam`` over the (decay / no-decay)
    parameter groups + ``ReduceLROnPlateau(optimizer, 'max')`` -- construction and the scheduler's
    view of the groups need no device; a torch module's parameters are refused (no per-tensor
    fallback)."""
    m = mmu.MIMOTransfomer(out_dim=4, num_classes=10, image_dim=196, hidden_size=64,
                           multimodal_num_attention_heads=2, multimodal_num_hidden_layers=2, drop=0.0)
    named = list(m.named_parameters())
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
              {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    assert groups[0]["params"] and groups[1]["params"]
    opt = mmu.BertAdam(groups, lr=1e-3, warmup=0.1, t_total=100)
    assert [g["weight_decay"] for g in opt.param_groups] == [0.01, 0.0] and len(opt._owners) == 1
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "max", patience=0, factor=0.5)
    sched.step(50.0)
    sched.step(40.0)                                    # worse: halves the lr of both groups
    assert [g["lr"] for g in opt.param_groups] == [5e-4, 5e-4]
    with pytest.raises(ValueError):
        mmu.BertAdam(torch.nn.Linear(2, 2).parameters(), lr=1e-3)
    with pytest.raises(ValueError):
        mmu.BertAdam(groups, lr=1e-3, schedule="warmup_cosine")
