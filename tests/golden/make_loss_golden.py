"""Golden vectors for the photometric loss, produced by IMPORTING the reference's own
frosting_utils/loss_utils.py (it needs only torch, so it runs in the build container on CPU):

    python tests/golden/make_loss_golden.py        # writes tests/golden/loss_l1dssim.npz

Stores inputs, the reference's loss 0.8*l1 + 0.2*(1-ssim) (frosting_trainers/refine.py:407-409) and its autograd
gradient w.r.t. the prediction."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/frosting_utils/loss_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

g = torch.Generator().manual_seed(2024)
out = {}
for name, (H, W, noise) in {"a": (45, 70, 0.08), "b": (64, 48, 0.3), "c": (16, 16, 0.02)}.items():
    gt = torch.rand(1, 3, H, W, generator=g)
    pred = (gt + noise * torch.randn(1, 3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    loss = 0.8 * ref.l1_loss(pred, gt) + 0.2 * (1.0 - ref.ssim(pred, gt))
    loss.backward()
    out[f"{name}_pred"] = pred.detach().numpy()[0]
    out[f"{name}_gt"] = gt.numpy()[0]
    out[f"{name}_loss"] = np.float32(loss.item())
    out[f"{name}_grad"] = pred.grad.numpy()[0]
np.savez_compressed(os.path.join(HERE, "loss_l1dssim.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("loss")})
