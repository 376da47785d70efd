import sys, os, types, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import mmu_b200
from det_params import det_image_encoder_state
ie = importlib.import_module("multi-modal-uncertainty_b200.src.image_encoder")
c = torch.load(os.path.join(ROOT, "tests/golden/image_encoder.pt"), weights_only=False)["avg3"]
cfg = c["cfg"]
for layers, width in (((1, 2, 1, 1), 16), ((1, 1, 1, 1), 64), ((1, 1, 1, 1), 32), ((1,1,1,1), 16)):
    outs = {}
    for prec in ("fp32", "bf16"):
        args = types.SimpleNamespace(num_image_embeds=3, img_embed_pool_type="avg", precision=prec,
                                     img_encoder_layers=layers, img_encoder_width=width)
        enc = ie.ImageEncoder(args)
        enc.load_state_dict(det_image_encoder_state({k: v.shape for k, v in enc.state_dict().items()}, 41))
        enc = enc.cuda().eval()
        with torch.no_grad():
            outs[prec] = enc(c["x"].cuda()).cpu()
    a, b = outs["fp32"], outs["bf16"]
    print(layers, width, "max fp32", float(a.abs().max()), "max bf16", float(b.abs().max()),
          "rel", float((a - b).abs().max() / a.abs().max()))
