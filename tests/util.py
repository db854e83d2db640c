"""Shared helpers for the parity tests."""
import torch

from frosting_b200 import scenes


def scene(P, W, H, seed, sh_degree, device, bg=0.0):
    cam = scenes.make_camera(W, H, device=device)
    g = scenes.random_gaussians(P, cam, seed, device=device)
    rs = scenes.settings_for(cam, sh_degree, bg=torch.full((3,), float(bg)), device=device)
    return cam, g, rs


def rel_err_stats(a, b):
    """Gradient tolerance used throughout: error relative to the tensor's scale, plus the fraction of
    significant elements whose own relative error exceeds 1e-3."""
    a, b = a.double().flatten(), b.double().flatten()
    scale = b.abs().max().clamp_min(1e-30)
    diff = (a - b).abs()
    max_rel_to_scale = (diff.max() / scale).item()
    sig = b.abs() > 1e-4 * scale
    if sig.any():
        el = diff[sig] / b.abs()[sig]
        frac_bad = (el > 1e-3).double().mean().item()
    else:
        frac_bad = 0.0
    return max_rel_to_scale, frac_bad
