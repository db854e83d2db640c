"""Dev helper: where a data-parallel training iteration spends its device time (frame vs step), under torchrun."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from frosting_b200 import camera_batch as cb

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
wl = cb.build_workload("c3", dev, rank, world)
step = bench.OursDPTrain(wl, dev)
cams = wl["cams"]
H, W = wl["H"], wl["W"]
step.gt = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(99 + rank)).to(dev) for _ in cams]
rs = [step.settings(c) for c in cams]
cot = [c.to(dev) for c in wl["cot_host"]]
ev = lambda: torch.cuda.Event(enable_timing=True)
import time
tf, ts = [], []
# wall clock of a free-running loop (what bench.py measures) next to the device times of its parts
for phase in range(2):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for it in range(16):
        i = it % len(cams)
        loss = cb.CameraBatch.frame(step, i, rs[i], cot[i]).reshape(1).clone()
        step.opt.update_learning_rate()
        step.opt.step(loss=loss)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize(dev)
    t_all = time.perf_counter() - t0
    if rank == 0:
        print(f"free-running 16 iterations: host issue {t_host / 16 * 1e3:.3f} ms/iter, wall {t_all / 16 * 1e3:.3f} ms/iter")
for it in range(12):
    i = it % len(cams)
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    loss = cb.CameraBatch.frame(step, i, rs[i], cot[i]).reshape(1).clone()
    e1.record()
    step.opt.update_learning_rate()
    step.opt.step(loss=loss)
    e2.record()
    torch.cuda.synchronize(dev)
    tf.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2))
if rank == 0:
    print(f"world {world} {step.opt.slabs.transport}: frame ms {[round(x, 2) for x in tf[2:]]}")
    print(f"step ms {[round(x, 2) for x in ts[2:]]}")
step.opt.close()
if world > 1:
    dist.destroy_process_group()
