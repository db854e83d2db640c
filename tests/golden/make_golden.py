"""Generates tests/golden/*.npz with the UNMODIFIED reference (oracle/_ref) -- run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   then copy the .npz files to tests/golden/

Each fixture stores the INPUTS as well (torch's CPU RNG is not bit-stable across CPU types), the reference's
outputs and the internal buffers the parity tests compare.  Small on purpose (< 1 MB each)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from frosting_b200 import scenes      # noqa: E402
from oracle import refdgr             # noqa: E402

CASES = {
    # name: (P, W, H, seed, sh_degree, bg, mode)
    "sh3_small": (1500, 112, 80, 101, 3, 0.0, "sh"),
    "sh0_bg1": (1200, 96, 96, 102, 0, 1.0, "sh"),
    "precomp": (1000, 80, 64, 103, 0, 0.5, "precomp"),
}


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, (P, W, H, seed, D, bg, mode) in CASES.items():
        cam = scenes.make_camera(W, H, device=dev)
        g = scenes.random_gaussians(P, cam, seed, device=dev, large_frac=0.01)
        rs = scenes.settings_for(cam, D, bg=torch.full((3,), bg), device=dev)
        gen = torch.Generator().manual_seed(seed + 1000)
        cot = torch.randn(3, H, W, generator=gen).to(dev)
        kw = dict(scales=g["scales"], rotations=g["rotations"])
        if mode == "sh":
            kw["shs"] = g["shs"]
        else:
            kw["colors_precomp"] = torch.rand(P, 3, generator=gen).to(dev)
        ref = refdgr.forward(rs, g["means3D"], g["opacities"], **kw)
        bwd = refdgr.backward(rs, ref, g["means3D"], cot, **kw)
        R = ref["num_rendered"]
        gv, bv, iv = refdgr.geom_views(ref["geom"], P), refdgr.binning_views(ref["binning"], R), refdgr.img_views(ref["img"], H, W)
        c = lambda t: t.detach().cpu().numpy()
        out = dict(
            P=P, W=W, H=H, D=D, bg=np.float32(bg), tanfovx=np.float64(rs.tanfovx), tanfovy=np.float64(rs.tanfovy),
            viewmatrix=c(rs.viewmatrix), projmatrix=c(rs.projmatrix), campos=c(rs.campos),
            means3D=c(g["means3D"]), opacities=c(g["opacities"]), scales=c(g["scales"]), rotations=c(g["rotations"]),
            cot=c(cot), color=c(ref["color"]), radii=c(ref["radii"]), num_rendered=R,
            depths=c(gv["depths"]), means2D=c(gv["means2D"]), conic_opacity=c(gv["conic_opacity"]), rgb=c(gv["rgb"]),
            tiles_touched=c(gv["tiles_touched"]), point_list=c(bv["point_list"]), ranges=c(iv["ranges"]),
            n_contrib=c(iv["n_contrib"]), final_T=c(iv["accum_alpha"]),
            **{"g_" + k: c(v) for k, v in bwd.items()})
        if mode == "sh":
            out["shs"] = c(g["shs"])
        else:
            out["colors_precomp"] = c(kw["colors_precomp"])
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "R =", R, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
