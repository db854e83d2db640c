"""Run W warm-up + N measured fwd+bwd frames of one scene; meant to sit under ncu.
Usage: python tools/prof_once.py P W H D [n] [impl=mine|ref]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import frosting_b200 as fb
from frosting_b200 import scenes
from oracle import refdgr

P, W, H, D = (int(x) for x in sys.argv[1:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 1
impl = sys.argv[6] if len(sys.argv) > 6 else "mine"
dev = torch.device("cuda:0")
cam = scenes.make_camera(W, H, device=dev)
g = scenes.random_gaussians(P, cam, 1234, device=dev)
rs = scenes.settings_for(cam, D, device=dev)
cot = torch.randn(3, H, W, device=dev)
kw = dict(shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
leaves = {k: g[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
for it in range(n):
    if impl == "mine":
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        c, _ = fb.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"],
                                         shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        c.backward(cot)
    else:
        f = refdgr.forward(rs, g["means3D"], g["opacities"], **kw)
        refdgr.backward(rs, f, g["means3D"], cot, **kw)
torch.cuda.synchronize()
print("done")
