import sys, os, types, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import mmu_b200
from det_params import det_image_encoder_state
ie = importlib.import_module("multi-modal-uncertainty_b200.src.image_encoder")
c = torch.load(os.path.join(ROOT, "tests/golden/image_encoder.pt"), weights_only=False)["avg3"]
cfg = c["cfg"]
grads = {}
for prec in ("fp32", "bf16"):
    args = types.SimpleNamespace(num_image_embeds=cfg["n_img"], img_embed_pool_type=cfg["pool"], precision=prec,
                                 img_encoder_layers=tuple(cfg["layers"]), img_encoder_width=cfg["width"])
    enc = ie.ImageEncoder(args)
    enc.load_state_dict(det_image_encoder_state(c["state_dict_shapes"], cfg["seed"]))
    enc = enc.cuda().train()
    enc.zero_grad()
    tok = enc(c["x"].cuda())
    (tok * c["r"].cuda()).sum().backward()
    grads[prec] = {k: p.grad.detach().double().cpu().flatten().clone() for k, p in enc.named_parameters()}
for k in grads["fp32"]:
    a, b = grads["fp32"][k], grads["bf16"][k]
    cos = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
    print(f"{k:40s} cos {cos:.4f}  norm ratio {float(b.norm() / a.norm()):.3f}")
