"""Generate golden fixtures by running the UNMODIFIED reference modules (container-only).

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/*.pt

The reference cannot travel to the GPU box, so its outputs on small seeded inputs are frozen
here.  ``tests/test_oracle_golden.py`` pins the oracle to them; the GPU parity tests then
compare the CUDA path with the oracle (and with these fixtures directly).

Import recipe: DATA_DIR must be set at import time; ``matplotlib`` and
``pytorch_pretrained_bert`` are absent and only imported, never used on this path, so empty
stub modules stand in for them.  Nothing of the reference is copied into the repo.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    os.environ.setdefault("DATA_DIR", "/tmp/mmu_data")
    os.environ.setdefault("RESULTS_DIR", "/tmp/mmu_results")
    for name in ["matplotlib", "matplotlib.style", "matplotlib.pyplot", "pytorch_pretrained_bert",
                 "pytorch_pretrained_bert.modeling"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].style = sys.modules["matplotlib.style"]

    class _Absent:
        def __init__(self, *a, **k):
            raise RuntimeError("stub for an absent third-party dependency")

    sys.modules["pytorch_pretrained_bert"].BertTokenizer = _Absent
    sys.modules["pytorch_pretrained_bert"].BertAdam = _Absent
    sys.modules["pytorch_pretrained_bert.modeling"].BertModel = _Absent
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.dataset as ref_dataset
    import src.framework as ref_framework
    import src.model as ref_model
    return ref_model, ref_dataset, ref_framework


def det_state_dict(shapes, seed):
    """Deterministic parameters independent of any library's init code: N(0, 0.05) for
    matrices, small offsets for vectors, LayerNorm weights near 1."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        t = torch.randn(*shape, generator=g)
        if name.endswith("weight") and len(shape) == 1:  # LayerNorm gain
            sd[name] = 1.0 + 0.1 * t
        elif len(shape) == 1:
            sd[name] = 0.1 * t
        elif name == "class_embeddings":
            sd[name] = shape[0] ** -0.5 * t
        else:
            sd[name] = t * (1.0 / np.sqrt(shape[-1]))
    return sd


def ref_acc():
    """``acc`` lives in train.py, which cannot be imported (argparse/__main__ globals); pull the
    function's source out of the file and exec it unchanged."""
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "acc"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train.py:acc", "exec"), ns)
    return ns["acc"]


def flava_case(ref_model, name, *, cls, avg_pool, E, C, D, heads, layers, d_img, d_txt, l_img,
               l_txt, B, seed):
    klass = ref_model.FlavaFusionTransfomerwithCLSToken if cls else ref_model.FlavaFusionTransfomer
    model = klass(out_dim=E, num_classes=C, image_hidden_size=d_img, text_hidden_size=d_txt,
                  multimodal_hidden_size=D, multimodal_num_attention_heads=heads,
                  multimodal_num_hidden_layers=layers, drop=0.0, avg_pool=avg_pool)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = det_state_dict(shapes, seed)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn(B, l_img, d_img, generator=g)
    txt = torch.randn(B, l_txt, d_txt, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    y_train = y.unsqueeze(1).repeat(1, E)
    acc = ref_acc()

    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1.0e-9,
                            weight_decay=1e-3)
    opt.zero_grad()
    logits = model((img, txt))
    loss = model.compute_loss(logits, y_train)
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in model.named_parameters()}
    train_acc = acc(logits.detach(), y_train, False, True)
    opt.step()
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}

    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        logits_eval = model((img, txt))
        loss_eval = model.compute_loss(logits_eval, y, eval=True)
        eval_acc = acc(logits_eval, y, True, True)
        extra = {}
        if cls:  # only the CLS variant can run a missing modality as committed
            extra["logits_img_only"] = model((img, None))
            extra["logits_txt_only"] = model((None, txt))
    out = dict(name=name, cfg=dict(cls=cls, avg_pool=avg_pool, E=E, C=C, D=D, heads=heads,
                                   layers=layers, d_img=d_img, d_txt=d_txt, l_img=l_img,
                                   l_txt=l_txt, B=B, seed=seed),
               state_dict=sd, img=img, txt=txt, y=y, y_train=y_train,
               logits=logits.detach(), loss=loss.detach(), grads=grads,
               train_acc=train_acc, params_after_adamw=after,
               logits_eval=logits_eval, loss_eval=loss_eval, eval_acc=eval_acc, **extra)
    return out


def big_case(ref_model):
    """Full-width model (D=768, 3 heads, 3 layers, E=5, C=101) on a small batch; only outputs and
    per-parameter gradient summaries are stored (the parameters are regenerated from the seed)."""
    E, C, D, B, l_img, l_txt, seed = 5, 101, 768, 8, 9, 5, 7
    model = ref_model.FlavaFusionTransfomer(out_dim=E, num_classes=C, avg_pool=False)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = det_state_dict(shapes, seed)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn(B, l_img, D, generator=g)
    txt = torch.randn(B, l_txt, D, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    y_train = y.unsqueeze(1).repeat(1, E)
    model.train()
    logits = model((img, txt))
    loss = model.compute_loss(logits, y_train)
    loss.backward()
    gsum = {k: torch.stack([p.grad.double().sum(), p.grad.double().abs().sum(),
                            p.grad.double().pow(2).sum()])
            if p.grad is not None else torch.zeros(3, dtype=torch.float64)
            for k, p in model.named_parameters()}
    gslice = {k: p.grad.reshape(-1)[:64].clone() for k, p in model.named_parameters()
              if p.grad is not None}
    return dict(cfg=dict(E=E, C=C, D=D, heads=3, layers=3, B=B, l_img=l_img, l_txt=l_txt,
                         seed=seed, avg_pool=False, cls=False, d_img=D, d_txt=D),
                shapes=shapes, img=img, txt=txt, y=y, y_train=y_train, logits=logits.detach(),
                loss=loss.detach(), grad_summaries=gsum, grad_slices=gslice)


def mimo_transformer_case(ref_model):
    E, C, D, B, seed = 4, 10, 48, 6, 11
    model = ref_model.MIMOTransfomer(out_dim=E, num_classes=C, hidden_size=D, image_dim=196,
                                     multimodal_num_hidden_layers=2,
                                     multimodal_num_attention_heads=2, drop=0)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = det_state_dict(shapes, seed)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, 4, 1, 14, 14, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    y_train = y.unsqueeze(1).repeat(1, E)
    model.train()
    logits = model(x)
    loss = model.compute_loss(logits, y_train)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    return dict(cfg=dict(E=E, C=C, D=D, heads=2, layers=2, B=B, seed=seed), state_dict=sd, x=x,
                y=y, y_train=y_train, logits=logits.detach(), loss=loss.detach(), grads=grads)


def mimo_resnet_case(ref_model):
    """MIMOResNet (reference src/model.py:17-100, src/layers.py:7-38) on a (B, 4, 1, 14, 14) batch:
    train-mode logits / loss / every gradient / BatchNorm running statistics after the step's
    forward, then eval-mode logits with those statistics."""
    E, C, B = 4, 10, 4
    model = ref_model.MIMOResNet(num_channels=1, emb_dim=4, out_dim=E, num_classes=C)
    # Gradients through ReLU are only comparable between two fp32 implementations when no
    # pre-activation sits within rounding (~3e-6) of zero -- such an element flips its mask and,
    # with a handful of samples, moves whole BatchNorm gradients by percents.  Search for a seed
    # whose smallest |pre-activation| clears 2e-5.
    margins = []
    hooks = [m.register_forward_pre_hook(lambda mod, inp: margins.append(float(inp[0].detach().abs().min())))
             for m in model.modules() if isinstance(m, torch.nn.ReLU)]
    seed = 31
    while True:
        case = _mimo_resnet_at_seed(model, seed, E, C, B, margins)
        if case is not None:
            break
        seed += 1
    for h in hooks:
        h.remove()
    return case


def _mimo_resnet_at_seed(model, seed, E, C, B, margins):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros_like(v)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif v.dim() == 4:   # conv (O, I, kh, kw)
            sd[k] = torch.randn(v.shape, generator=g) / np.sqrt(v.shape[1] * v.shape[2] * v.shape[3])
        elif v.dim() == 2:   # fc
            sd[k] = torch.randn(v.shape, generator=g) / np.sqrt(v.shape[1])
        elif k.endswith("weight"):   # BN gain
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:                # BN / fc bias
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
    model.load_state_dict(sd, strict=True)
    x = torch.rand(B, 4, 1, 14, 14, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    y_train = y.unsqueeze(1).repeat(1, E)
    model.train()
    model.zero_grad()
    del margins[:]
    logits = model(x)
    if min(margins) < 2e-5:
        return None
    train_margin = min(margins)
    loss = model.compute_loss(logits, y_train)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
    return dict(cfg=dict(E=E, C=C, B=B, seed=seed, num_channels=1, emb_dim=4, relu_margin=train_margin),
                state_dict=sd, x=x,
                y=y, y_train=y_train, logits=logits.detach(), loss=loss.detach(), grads=grads,
                state_after_forward=after, logits_eval=logits_eval,
                param_order=[k for k, _ in model.named_parameters()],
                buffer_order=[k for k, _ in model.named_buffers()])


def shaping_cases(ref_dataset):
    out = {}
    g = torch.Generator().manual_seed(3)
    img = torch.randn(6, 4, 8, generator=g)
    txt = torch.randn(6, 3, 8, generator=g)
    y = torch.randint(0, 5, (6,), generator=g)
    for mt in ["Vanilla", "MultiHead", "MIMO-shuffle-instance"]:
        for phase in ["train", "eval"]:
            torch.manual_seed(42)
            (i2, t2), y2 = ref_dataset.data_forming_func_transformer((img, txt), y, phase, mt)
            out[f"transformer/{mt}/{phase}"] = dict(img=i2.clone(), txt=t2.clone(), y=y2.clone())
    x = torch.rand(5, 4, 1, 14, 14, generator=g)
    yv = torch.randint(0, 10, (5,), generator=g)
    for mt in ["Vanilla", "single-model-weight-sharing", "MultiHead", "MIMO-shuffle-instance",
               "MIMO-shuffle-view", "MIMO-shuffle-all"]:
        for phase in ["train", "eval"]:
            torch.manual_seed(42)
            x2, y2 = ref_dataset.data_forming_func(x, yv, phase, mt)
            out[f"fmnist/{mt}/{phase}"] = dict(x=x2.clone(), y=y2.clone())
    out["inputs"] = dict(img=img, txt=txt, y=y, x=x, yv=yv)
    ragged = [(torch.randn(l1, 8, generator=g), torch.randn(l2, 8, generator=g),
               torch.LongTensor([c])) for l1, l2, c in [(4, 2, 1), (2, 5, 0), (3, 3, 4)]]
    (pi, pt), pl = ref_dataset.collate_fn_flava(ragged)
    out["collate"] = dict(ragged=ragged, img=pi, txt=pt, labels=pl)
    return out


def fmnist_views_case(ref_dataset):
    """Either side of the FashionMNIST path: (1) the quarter-view input format -- the reference's
    ``QuarterCrop((28, 28))`` + per-crop ``ToTensor`` stack exactly as ``get_fmnist`` composes them
    (src/dataset.py:105-151) on seeded uint8 images; (2) the four-view robustness sweep of
    ``eval_robustness.py:82-122``: the statements of the script's own ``__main__`` body, from
    ``outputs = []`` to ``labels = np.concatenate(...)``, are executed from its AST (the script
    itself cannot be imported: argparse, dataset download, checkpoint) with a deterministic linear
    stand-in for the model -- once for a multi-head model (view i zero-filled) and once for
    ``single-model-weight-sharing`` (view i removed, the remaining three views through the shared
    model)."""
    from PIL import Image
    from torchvision import transforms
    rng = np.random.RandomState(5)
    imgs = rng.randint(0, 256, size=(6, 28, 28)).astype(np.uint8)
    tq = transforms.Compose([
        ref_dataset.QuarterCrop((28, 28)),
        transforms.Lambda(lambda crops: torch.stack([transforms.ToTensor()(crop) for crop in crops]))])
    quarters = torch.stack([tq(Image.fromarray(im, mode="L")) for im in imgs])      # (6, 4, 1, 14, 14)

    src = open(os.path.join(REF, "eval_robustness.py")).read()
    tree = ast.parse(src)
    main_if = [n for n in tree.body if isinstance(n, ast.If)][-1]
    names = [ast.unparse(n.targets[0]) if isinstance(n, ast.Assign) else None for n in main_if.body]
    first = names.index("outputs")
    last = max(i for i, n in enumerate(names) if n == "labels")
    body = [n for n in main_if.body[first:last + 1]
            if not (isinstance(n, ast.Expr) and isinstance(n.value, ast.Call)
                    and getattr(n.value.func, "id", "") == "print")]
    code = compile(ast.Module(body=body, type_ignores=[]), "eval_robustness_main", "exec")

    g = torch.Generator().manual_seed(9)
    C_ = 10
    valid = [(torch.rand(b, 4, 1, 14, 14, generator=g), torch.randint(0, C_, (b,), generator=g)) for b in (5, 3)]
    W4 = torch.randn(4 * 196, 4 * C_, generator=g)      # multi-head stand-in: (B, 4, 1, 14, 14) -> (B, 4, C)
    W1 = torch.randn(196, C_, generator=g)               # weight sharing: (B * 3, 1, 14, 14) -> (B * 3, 1, C)
    out = {"images_u8": torch.from_numpy(imgs), "quarters": quarters, "valid": valid, "W4": W4, "W1": W1}
    for model_type, model in (("MultiHead", lambda x: (x.reshape(x.shape[0], -1) @ W4).view(-1, 4, C_)),
                              ("single-model-weight-sharing",
                               lambda x: (x.reshape(x.shape[0], -1) @ W1).view(-1, 1, C_))):
        ns = {"torch": torch, "np": np, "dataset": ref_dataset, "model": model, "valid": valid,
              "args": types.SimpleNamespace(model_type=model_type, device="cpu")}
        exec(code, ns)
        out[model_type] = dict(outputs=torch.from_numpy(ns["outputs"]), labels=torch.from_numpy(ns["labels"]))

    # (3) the plain prediction dump of eval_prediction_saving.py:77-104, same recipe
    src = open(os.path.join(REF, "eval_prediction_saving.py")).read()
    main_if = [n for n in ast.parse(src).body if isinstance(n, ast.If)][-1]
    names = [ast.unparse(n.targets[0]) if isinstance(n, ast.Assign) else None for n in main_if.body]
    first = names.index("outputs")
    last = max(i for i, n in enumerate(names) if n == "labels")
    body = [n for n in main_if.body[first:last + 1]
            if not (isinstance(n, ast.Expr) and isinstance(n.value, ast.Call)
                    and getattr(n.value.func, "id", "") == "print")]
    code = compile(ast.Module(body=body, type_ignores=[]), "eval_prediction_saving_main", "exec")
    for model_type, model in (("MultiHead", lambda x: (x.reshape(x.shape[0], -1) @ W4).view(-1, 4, C_)),
                              ("single-model-weight-sharing",
                               lambda x: (x.reshape(x.shape[0], -1) @ W1).view(-1, 1, C_))):
        ns = {"torch": torch, "np": np, "dataset": ref_dataset, "model": model, "valid": valid,
              "args": types.SimpleNamespace(model_type=model_type, device="cpu")}
        exec(code, ns)
        out["predictions/" + model_type] = dict(outputs=torch.from_numpy(ns["outputs"]),
                                                labels=torch.from_numpy(ns["labels"]))
    return out


def mmbt_collate_case(ref_dataset):
    """The MMBT input format (reference src/dataset.py:371-438): rows as ``JsonlDataset.__getitem__``
    returns them -- (token ids (l_i,) int64, segment (l_i,) FLOAT ones, image (3, h, w), label (1,)) --
    through the reference's own ``collate_fn``."""
    g = torch.Generator().manual_seed(17)
    rows = []
    for l, c in ((5, 1), (2, 0), (7, 1), (1, 0)):
        rows.append((torch.randint(1, 30000, (l,), generator=g), torch.zeros(l) + 1,
                     torch.randn(3, 4, 4, generator=g), torch.LongTensor([c])))
    (txt, segment, mask, img), tgt = ref_dataset.collate_fn(rows)
    return dict(rows=rows, txt=txt, segment=segment, mask=mask, img=img, tgt=tgt)


def sampling_case():
    """``input_sampling`` lives in a script that cannot be imported (argparse at module level is
    under __main__, but it imports dataset loaders that do not exist); exec its source."""
    src = open(os.path.join(REF, "eval_transformer_robustness.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "input_sampling"][0]
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "input_sampling", "exec"), ns)
    np.random.seed(42)
    torch.manual_seed(42)
    draws = []
    for type in ["image", "text"]:
        for _ in range(20):
            ii, it = ns["input_sampling"](197, 40, type)
            draws.append((ii.clone(), it.clone()))
    return dict(l_img=197, l_txt=40, n_repeats=20, np_seed=42, torch_seed=42, draws=draws)


def optimizer_case():
    from transformers.optimization import get_cosine_schedule_with_warmup
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(257, generator=g)
    grads = [torch.randn(257, generator=g) * (0.1 + i) for i in range(6)]
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=3e-3, betas=(0.9, 0.98), eps=1.0e-9, weight_decay=1e-3)
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=3, num_training_steps=10)
    traj, lrs = [], []
    for gr in grads:
        lrs.append(opt.param_groups[0]["lr"])
        p.grad = gr.clone()
        opt.step()
        sched.step()
        traj.append(p.detach().clone())
    st = opt.state[p]
    return dict(p0=p0, grads=grads, lr=3e-3, warmup=3, total=10, lrs=lrs, traj=traj,
                exp_avg=st["exp_avg"].clone(), exp_avg_sq=st["exp_avg_sq"].clone())


def notebook_case():
    src = open(os.path.join(REF, "notebooks", "utils.py")).read()
    tree = ast.parse(src)
    from scipy.stats import pearsonr
    ns = {"np": np, "pearsonr": pearsonr}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef)
           and n.name in ("softmax", "get_correlation")]
    exec(compile(ast.Module(body=fns, type_ignores=[]), "nb_utils", "exec"), ns)
    src2 = open(os.path.join(REF, "notebooks", "food101_robustness.py")).read()
    tree2 = ast.parse(src2)
    fns2 = [n for n in tree2.body if isinstance(n, ast.FunctionDef)
            and n.name == "process_predictions_food101"]
    exec(compile(ast.Module(body=fns2, type_ignores=[]), "nb_food", "exec"), ns)
    rng = np.random.RandomState(9)
    S, V, K, C = 50, 43, 2, 6
    preds = rng.randn(S, V, K, C).astype(np.float32) * 2
    labels = rng.randint(0, C, size=S)
    lab, ori, image, text, ic, tc = ns["process_predictions_food101"](preds, labels)
    corr = ns["get_correlation"](lab, ori, image, text, ic, tc)
    pred = preds.mean(2).argmax(-1)
    return dict(preds=torch.from_numpy(preds), labels=torch.from_numpy(labels),
                ori=torch.from_numpy(ori), image=torch.from_numpy(image),
                text=torch.from_numpy(text), image_corr=torch.from_numpy(ic),
                text_corr=torch.from_numpy(tc), corr_image=float(corr["image"]),
                corr_text=float(corr["text"]),
                acc_full=float((pred[:, 0] == labels).mean() * 100))


def rank_case():
    """Rank statistics of the notebooks, computed by the reference's own functions:
    `trunk_pred_top` / `subnetwork_wise_kendalltau` (notebooks/analysis_round_1.py:74-90) and
    `process_predictions_hatefulmeme` / `AUC_table` (notebooks/hatefulmeme_robustness.py:22-41,
    105-112), extracted from the unmodified sources with ast and executed on seeded arrays."""
    import itertools
    import pandas as pd
    import scipy.stats as stats
    from sklearn.metrics import roc_auc_score
    ns = {"np": np, "stats": stats, "itertools": itertools}
    tree = ast.parse(open(os.path.join(REF, "notebooks", "analysis_round_1.py")).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef)
           and n.name in ("trunk_pred_top", "subnetwork_wise_kendalltau")]
    exec(compile(ast.Module(body=fns, type_ignores=[]), "nb_round1", "exec"), ns)
    rng = np.random.RandomState(31)
    S, E, C, top = 300, 4, 10, 5
    predictions = rng.randn(S, E, C).astype(np.float32)
    predictions[::7, :, 3] = predictions[::7, :, 5]          # ties inside rows
    predictions[5] = np.round(predictions[5])                # many equal entries
    labels = rng.randint(0, C, size=S)
    muted = [ns["trunk_pred_top"](predictions[:, i, :], labels, top, mute_true=True)
             for i in range(E)]
    plain = ns["trunk_pred_top"](predictions[:, 0, :], labels, 3, mute_true=False)
    taus = ns["subnetwork_wise_kendalltau"](muted)

    ns2 = {"np": np, "pd": pd, "roc_auc_score": roc_auc_score}
    src_u = open(os.path.join(REF, "notebooks", "utils.py")).read()
    fn_u = [n for n in ast.parse(src_u).body if isinstance(n, ast.FunctionDef) and n.name == "softmax"]
    exec(compile(ast.Module(body=fn_u, type_ignores=[]), "nb_utils", "exec"), ns2)
    tree2 = ast.parse(open(os.path.join(REF, "notebooks", "hatefulmeme_robustness.py")).read())
    fns2 = [n for n in tree2.body if isinstance(n, ast.FunctionDef)
            and n.name in ("AUC_table", "process_predictions_hatefulmeme")]
    exec(compile(ast.Module(body=fns2, type_ignores=[]), "nb_hateful", "exec"), ns2)
    S2, V, K = 400, 43, 2
    preds = (rng.randn(S2, V, K, 2) * 1.5).astype(np.float32)
    preds[::5] = np.round(preds[::5])                        # tied scores across samples
    lab2 = rng.randint(0, 2, size=S2)
    args = ns2["process_predictions_hatefulmeme"](preds, lab2)
    df = ns2["AUC_table"](*args)
    return dict(predictions=torch.from_numpy(predictions), labels=torch.from_numpy(labels), top=top,
                muted=torch.from_numpy(np.stack(muted).astype(np.float64)),
                plain_top3=torch.from_numpy(plain.astype(np.float64)),
                taus=torch.from_numpy(np.asarray(taus, dtype=np.float64)),
                hm_preds=torch.from_numpy(preds), hm_labels=torch.from_numpy(lab2),
                hm_scores=torch.from_numpy(np.concatenate(
                    [np.stack(args[1:4], 1), args[4], args[5]], 1).astype(np.float32)),
                hm_auc=torch.from_numpy(df["AUC"].to_numpy(dtype=np.float64)))


def init_case(ref_model):
    """Seeded construction of the reference modules: per-parameter checksums of the initial
    weights (the product constructs torch's own layers in the same order under the same seed)."""
    out = {}
    for name, klass, kw in [
            ("plain", ref_model.FlavaFusionTransfomer, dict(out_dim=2, num_classes=7, avg_pool=False)),
            ("cls", ref_model.FlavaFusionTransfomerwithCLSToken, dict(out_dim=3, num_classes=5, avg_pool=False))]:
        torch.manual_seed(123)
        m = klass(image_hidden_size=32, text_hidden_size=48, multimodal_hidden_size=64,
                  multimodal_num_attention_heads=2, multimodal_num_hidden_layers=2, drop=0.0, **kw)
        out[name] = {k: torch.stack([v.double().sum(), v.double().abs().sum()])
                     for k, v in m.state_dict().items()}
        out[name + "_param_order"] = [k for k, _ in m.named_parameters()]
    torch.manual_seed(123)
    m = ref_model.MIMOTransfomer(out_dim=4, num_classes=10, hidden_size=48,
                                 multimodal_num_hidden_layers=2, multimodal_num_attention_heads=2)
    out["mimo"] = {k: torch.stack([v.double().sum(), v.double().abs().sum()])
                   for k, v in m.state_dict().items()}
    out["mimo_param_order"] = [k for k, _ in m.named_parameters()]
    torch.manual_seed(123)
    m = ref_model.MIMOResNet(num_channels=1, emb_dim=4, out_dim=4, num_classes=10)
    out["resnet"] = {k: torch.stack([v.double().sum(), v.double().abs().sum()])
                     for k, v in m.state_dict().items()}
    out["resnet_param_order"] = [k for k, _ in m.named_parameters()]
    return out


def main():
    torch.set_num_threads(4)
    if sys.argv[1:] == ["rank"]:   # only the rank-statistics fixture (no model import needed)
        torch.save(rank_case(), os.path.join(HERE, "rank_stats.pt"))
        return
    ref_model, ref_dataset, _ = import_reference()
    if sys.argv[1:] == ["mmbt_collate"]:
        torch.save(mmbt_collate_case(ref_dataset), os.path.join(HERE, "mmbt_collate.pt"))
        return
    if sys.argv[1:] == ["fmnist_views"]:   # only the FashionMNIST view-format / view-sweep fixture
        torch.save(fmnist_views_case(ref_dataset), os.path.join(HERE, "fmnist_views.pt"))
        return
    small = dict(C=7, D=64, heads=2, layers=2, d_img=32, d_txt=48, l_img=5, l_txt=3, B=4)
    cases = {
        "plain_E2": flava_case(ref_model, "plain_E2", cls=False, avg_pool=False, E=2, seed=21, **small),
        "plain_E1": flava_case(ref_model, "plain_E1", cls=False, avg_pool=False, E=1, seed=22, **small),
        "avgpool_E2": flava_case(ref_model, "avgpool_E2", cls=False, avg_pool=True, E=2, seed=23, **small),
        "cls_E3": flava_case(ref_model, "cls_E3", cls=True, avg_pool=False, E=3, seed=24, **small),
        # hd = 256 / D = 768-like head geometry at reduced width: 3 heads
        "plain_E5_h3": flava_case(ref_model, "plain_E5_h3", cls=False, avg_pool=False, E=5, C=11,
                                  D=96, heads=3, layers=3, d_img=96, d_txt=96, l_img=7, l_txt=4,
                                  B=9, seed=25),
    }
    torch.save(cases, os.path.join(HERE, "flava_small.pt"))
    torch.save(big_case(ref_model), os.path.join(HERE, "flava_768.pt"))
    torch.save(mimo_transformer_case(ref_model), os.path.join(HERE, "mimo_transformer.pt"))
    torch.save(mimo_resnet_case(ref_model), os.path.join(HERE, "mimo_resnet.pt"))
    torch.save(shaping_cases(ref_dataset), os.path.join(HERE, "shaping.pt"))
    torch.save(sampling_case(), os.path.join(HERE, "input_sampling.pt"))
    torch.save(optimizer_case(), os.path.join(HERE, "adamw_cosine.pt"))
    torch.save(notebook_case(), os.path.join(HERE, "notebook_scoring.pt"))
    torch.save(init_case(ref_model), os.path.join(HERE, "init_seed123.pt"))
    torch.save(rank_case(), os.path.join(HERE, "rank_stats.pt"))
    torch.save(fmnist_views_case(ref_dataset), os.path.join(HERE, "fmnist_views.pt"))
    torch.save(mmbt_collate_case(ref_dataset), os.path.join(HERE, "mmbt_collate.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
