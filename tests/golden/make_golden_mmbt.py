"""Golden fixtures for the MMBT path: the UNMODIFIED reference ``src/mmbt.py`` run in the build
container (needs /root/reference), on top of ``oracle/bert_restated.py`` standing in for the
absent, unpinned third-party ``pytorch_pretrained_bert`` (restated from its published 0.6.2
definitions), and torchvision's own ``resnet152`` with random weights (``pretrained=True`` needs
the network).

    python tests/golden/make_golden_mmbt.py      # writes tests/golden/mmbt_small.pt, bertadam.pt

What this pins: everything ``src/mmbt.py`` itself does (ImageEncoder pooling, ImageBertEmbeddings,
mask construction, the four forward variants incl. ``forward_control``'s index sampling, the
classifier and loss) bit-for-bit as the reference executes it.  What it cannot pin: the BERT
arithmetic of the real third-party package (parity unpinned).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import bert_restated  # noqa: E402


def import_reference_mmbt():
    import torchvision
    os.environ.setdefault("DATA_DIR", "/tmp/mmu_data")
    os.environ.setdefault("RESULTS_DIR", "/tmp/mmu_results")
    for name in ["matplotlib", "matplotlib.style", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].style = sys.modules["matplotlib.style"]
    pkg = types.ModuleType("pytorch_pretrained_bert")
    mod = types.ModuleType("pytorch_pretrained_bert.modeling")
    mod.BertModel = bert_restated.BertModel
    pkg.modeling = mod
    pkg.BertAdam = bert_restated.BertAdam
    pkg.BertTokenizer = object
    sys.modules["pytorch_pretrained_bert"] = pkg
    sys.modules["pytorch_pretrained_bert.modeling"] = mod
    real = torchvision.models.resnet152
    torchvision.models.resnet152 = lambda pretrained=False, **kw: real(weights=None)  # no network
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.mmbt as ref_mmbt
    return ref_mmbt


class Vocab:
    def __init__(self, cls_id, sep_id):
        self.stoi = {"[CLS]": cls_id, "[SEP]": sep_id, "[PAD]": 0}


def make_args(name, *, D, n_img, C, cls_id, sep_id, pool="avg"):
    return types.SimpleNamespace(bert_model=name, hidden_sz=D, img_hidden_sz=2048, num_image_embeds=n_img,
                                 img_embed_pool_type=pool, dropout=0.0, n_classes=C,
                                 vocab=Vocab(cls_id, sep_id))


def mmbt_case(ref_mmbt, *, seed, B, S_txt, n_img, D, heads, layers, d_ff, vocab, max_pos, C, img_hw):
    name = f"golden-{D}-{layers}"
    bert_restated.BertModel.CONFIGS[name] = bert_restated.BertConfig(
        vocab_size=vocab, hidden_size=D, num_hidden_layers=layers, num_attention_heads=heads,
        intermediate_size=d_ff, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
        max_position_embeddings=max_pos, type_vocab_size=2, initializer_range=0.05)
    cls_id, sep_id = 5, 6
    torch.manual_seed(seed)
    model = ref_mmbt.MultimodalBertClf(make_args(name, D=D, n_img=n_img, C=C, cls_id=cls_id, sep_id=sep_id))
    # default inits leave biases and LayerNorms at (0, 1): perturb them so every term is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.startswith("enc.img_encoder"):
                continue
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    txt = torch.randint(7, vocab, (B, S_txt), generator=g)
    lens = torch.randint(S_txt // 2, S_txt + 1, (B,), generator=g)
    lens[0] = S_txt
    mask = (torch.arange(S_txt)[None] < lens[:, None]).long()
    txt = txt * mask  # [PAD] = 0 beyond the sentence (src/dataset.py:400-410 pads with zeros)
    segment = mask.clone()  # segment = 1 on the sentence, 0 on padding (src/dataset.py:400-403)
    img = torch.randn(B, 3, img_hw, img_hw, generator=g)
    y = torch.randint(0, C, (B,), generator=g)

    out = {"cfg": dict(B=B, S_txt=S_txt, n_img=n_img, d_img=2048, D=D, n_head=heads, n_layers=layers,
                       d_ff=d_ff, vocab=vocab, max_pos=max_pos, n_types=2, C=C, cls_id=cls_id,
                       sep_id=sep_id),
           "txt": txt, "mask": mask, "segment": segment, "img": img, "y": y}
    model.eval()  # BatchNorm of the random ResNet-152 in inference mode: tokens independent of B
    with torch.no_grad():
        tokens = model.enc.img_encoder(img)
        out["img_tokens"] = tokens.clone()
        out["logits_full"] = model(txt, mask, segment, img)
        out["logits_img_only"] = model.forward_img_only(txt, mask, segment, img)
        out["logits_txt_only"] = model.forward_txt_only(txt, mask, segment, img)
        ctl = {}
        for modal in ("image", "text"):
            torch.manual_seed(1000 + len(ctl))
            state = torch.get_rng_state()
            logits = model.forward_control(txt, mask, segment, img, modal)
            torch.set_rng_state(state)  # re-draw what forward_control drew (src/mmbt.py:198-201)
            total = S_txt + n_img + 2
            num = n_img + 1 if modal == "image" else S_txt
            ind = torch.cat([torch.zeros(1, dtype=torch.long),
                             torch.sort(torch.randperm(total - 1)[:num] + 1)[0]])
            ctl[modal] = {"seed": 1000 + len(ctl), "indices": ind, "logits": logits}
        out["control"] = ctl
    # gradients (train mode for BERT -- dropout is 0 -- with the image encoder kept in eval so the
    # tokens are the ones recorded above), through the pooled image tokens
    model.train()
    model.enc.img_encoder.eval()
    model.zero_grad()
    tok_grad = {}
    def keep_token_grad(module, inputs, output):
        output.register_hook(lambda gr: tok_grad.__setitem__("g", gr.clone()))

    handle = model.enc.img_encoder.register_forward_hook(keep_token_grad)
    logits = model(txt, mask, segment, img)
    loss = model.compute_loss(logits, y)
    loss.backward()
    handle.remove()
    out["loss"] = loss.detach()
    out["logits_train"] = logits.detach()
    out["dimg_tokens"] = tok_grad["g"]
    sd = {k: v.detach().clone() for k, v in model.state_dict().items() if not k.startswith("enc.img_encoder")}
    out["state_dict"] = sd
    out["state_dict_keys_all"] = [k for k in model.state_dict().keys() if not k.startswith("enc.img_encoder")]
    out["named_parameters"] = [k for k, _ in model.named_parameters() if not k.startswith("enc.img_encoder")]
    out["grads"] = {k: p.grad.detach().clone() for k, p in model.named_parameters()
                    if not k.startswith("enc.img_encoder")}
    return out


def image_encoder_case(ref_mmbt, *, seed, layers, width, n_img, pool, B, hw):
    """The reference's ImageEncoder class, unmodified, over a thin torchvision Bottleneck ResNet
    (``resnet152`` itself is ResNet(Bottleneck, [3, 8, 36, 3]); same class, fewer / thinner blocks
    so that the fixture stays small).  Train-mode tokens + running statistics, eval-mode tokens and
    the gradients of sum(tokens * r)."""
    import torchvision
    from torchvision.models.resnet import Bottleneck, ResNet
    saved = torchvision.models.resnet152
    torchvision.models.resnet152 = lambda pretrained=False, **kw: ResNet(Bottleneck, list(layers),
                                                                         width_per_group=width)
    try:
        torch.manual_seed(seed)
        enc = ref_mmbt.ImageEncoder(types.SimpleNamespace(num_image_embeds=n_img, img_embed_pool_type=pool))
    finally:
        torchvision.models.resnet152 = saved
    from det_params import det_image_encoder_state, grad_digest
    enc.load_state_dict(det_image_encoder_state({k: v.shape for k, v in enc.state_dict().items()}, seed),
                        strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, 3, hw, hw, generator=g)
    r = torch.randn(B, n_img, 2048, generator=g)
    out = {"cfg": dict(layers=list(layers), width=width, n_img=n_img, pool=pool, B=B, hw=hw, seed=seed),
           "state_dict_shapes": {k: tuple(v.shape) for k, v in enc.state_dict().items()},
           "named_parameters": [k for k, _ in enc.named_parameters()], "x": x, "r": r}
    enc.eval()
    with torch.no_grad():
        out["tokens_eval"] = enc(x).clone()
    enc.train()
    enc.zero_grad()
    tok = enc(x)
    (tok * r).sum().backward()
    out["tokens_train"] = tok.detach().clone()
    out["grads"] = {k: grad_digest(p.grad) for k, p in enc.named_parameters()}
    out["buffers_after"] = {k: v.detach().clone() for k, v in enc.state_dict().items()
                            if "running" in k or "num_batches" in k}
    return out


def bertadam_case():
    """The reference's optimizer configuration for MMBT (train.py:136-147) on a few small tensors,
    several steps, including a tensor whose gradient norm exceeds max_grad_norm."""
    g = torch.Generator().manual_seed(77)
    shapes = {"a.weight": (24, 16), "a.bias": (24,), "LayerNorm.weight": (16,), "b.weight": (8, 24)}
    params = {k: torch.nn.Parameter(torch.randn(*s, generator=g) * 0.3) for k, s in shapes.items()}
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [{"params": [p for n, p in params.items() if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
              {"params": [p for n, p in params.items() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = bert_restated.BertAdam(groups, lr=5e-2, warmup=0.1, t_total=40)
    out = {"init": {k: p.detach().clone() for k, p in params.items()}, "grads": [], "after": [],
           "hyper": dict(lr=5e-2, warmup=0.1, t_total=40, b1=0.9, b2=0.999, e=1e-6, max_grad_norm=1.0),
           "decay": {k: (0.0 if any(nd in k for nd in no_decay) else 0.01) for k in shapes}}
    for step in range(6):
        gr = {k: torch.randn(*s, generator=g) * (3.0 if k == "a.weight" else 0.05) for k, s in shapes.items()}
        for k, p in params.items():
            p.grad = gr[k].clone()
        opt.step()
        out["grads"].append(gr)
        out["after"].append({k: p.detach().clone() for k, p in params.items()})
    return out


def main():
    torch.set_num_threads(4)
    ref_mmbt = import_reference_mmbt()
    cases = {} if "--only-image-encoder" in sys.argv else {
        "fp32_small": mmbt_case(ref_mmbt, seed=31, B=3, S_txt=11, n_img=3, D=128, heads=2, layers=2, d_ff=256,
                                vocab=120, max_pos=32, C=2, img_hw=64),
        # head_dim 64 (the tensor-core path's geometry), longer ragged text, 3 classes
        "hd64": mmbt_case(ref_mmbt, seed=32, B=4, S_txt=27, n_img=3, D=128, heads=2, layers=3, d_ff=512,
                          vocab=200, max_pos=64, C=3, img_hw=64),
    }
    if "--only-image-encoder" not in sys.argv:
        torch.save(cases, os.path.join(HERE, "mmbt_small.pt"))
        torch.save(bertadam_case(), os.path.join(HERE, "bertadam.pt"))
    enc_cases = {
        "avg3": image_encoder_case(ref_mmbt, seed=41, layers=(1, 2, 1, 1), width=16, n_img=3, pool="avg", B=4, hw=64),
        "max4": image_encoder_case(ref_mmbt, seed=43, layers=(1, 1, 1, 1), width=8, n_img=4, pool="max", B=3, hw=96),
    }
    torch.save(enc_cases, os.path.join(HERE, "image_encoder.pt"))
    for f in ("mmbt_small.pt", "bertadam.pt", "image_encoder.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
