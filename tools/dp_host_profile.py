"""Dev helper: cProfile of the host side of a training iteration (frosting_render + loss + backward + Adam step)."""
import cProfile
import os
import pstats
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from frosting_b200 import camera_batch as cb

dev = torch.device("cuda", 0)
wl = cb.build_workload("c3", dev, 0, 1)
step = bench.OursDPTrain(wl, dev)
cams = wl["cams"]
step.gt = [torch.rand(3, wl["H"], wl["W"]).to(dev) for _ in cams]
rs = [step.settings(c) for c in cams]
cot = [c.to(dev) for c in wl["cot_host"]]


def loop(n):
    for it in range(n):
        i = it % len(cams)
        loss = cb.CameraBatch.frame(step, i, rs[i], cot[i]).reshape(1).clone()
        step.opt.update_learning_rate()
        step.opt.step(loss=loss)


loop(10)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
loop(200)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
