"""Generates tests/golden/adam.npz from torch.optim.Adam (CPU) -- the optimizer the reference constructs at
frosting_scene/frosting_optimizer.py:101 -- with the reference's group learning rates and eps.  Inputs are stored in
the fixture, so it does not depend on torch's RNG staying stable.

    python tests/golden/make_adam_golden.py
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(4242)
    shapes = {"bary_coords": (37, 6), "sh_coordinates_dc": (37, 1, 3), "sh_coordinates_rest": (37, 15, 3),
              "opacities": (37, 1), "scales": (37, 3), "quaternions": (37, 4)}
    lrs = {"bary_coords": 0.005, "sh_coordinates_dc": 0.0025, "sh_coordinates_rest": 0.0025 / 20.0,
           "opacities": 0.05, "scales": 0.005, "quaternions": 0.001}
    steps = 6
    params = {n: torch.randn(s, generator=g).requires_grad_(True) for n, s in shapes.items()}
    out = {f"p0_{n}": p.detach().numpy().copy() for n, p in params.items()}
    opt = torch.optim.Adam([{"params": [params[n]], "lr": lrs[n], "name": n} for n in shapes], lr=0.0, eps=1e-15)
    for t in range(steps):
        for n, p in params.items():
            # a mix of magnitudes, exact zeros (occluded Gaussians get zero gradients) and sign changes
            gr = torch.randn(shapes[n], generator=g) * (10.0 ** torch.randint(-6, 1, shapes[n], generator=g).float())
            gr[torch.rand(shapes[n], generator=g) < 0.3] = 0.0
            p.grad = gr
            out[f"g{t}_{n}"] = gr.numpy().copy()
        opt.step()
        for n, p in params.items():
            out[f"p{t + 1}_{n}"] = p.detach().numpy().copy()
    for n in shapes:
        out[f"lr_{n}"] = np.float64(lrs[n])
    out["steps"] = np.int64(steps)
    out["names"] = np.array(list(shapes))
    np.savez_compressed(os.path.join(HERE, "adam.npz"), **out)
    print("wrote adam.npz")


if __name__ == "__main__":
    main()
