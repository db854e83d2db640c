"""Dev helper: device time of FrostingAdam.step() alone, dense vs sparse-row gradients, under torchrun.
    python -m torch.distributed.run --nproc-per-node N tools/dp_adam_probe.py [P]"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frosting_b200 import optim

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
g = torch.Generator(device=dev).manual_seed(1)
shapes = {"bary_logits": (P, 6), "sh_dc": (P, 1, 3), "sh_rest": (P, 15, 3), "opacity_logits": (P, 1), "log_scales": (P, 3), "quats": (P, 4)}
params = {k: torch.randn(s, device=dev, generator=g) for k, s in shapes.items()}
opt = optim.FrostingAdam.for_frosting(params)
radii = (torch.rand(P, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank)) < 0.1).int() * 5
# clustered like a real frame: runs of 4096 rows on / off
radii = ((torch.arange(P, device=dev) // 4096 + rank) % 10 == 0).int() * 5


def run(sparse, n=8):
    ts = []
    for it in range(n + 2):
        for k in shapes:
            opt.grads[k].normal_()
        if sparse:
            opt.grads.rows_from(radii, list(shapes))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.step()
        e1.record()
        torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


d, s = run(False), run(True)
if rank == 0:
    print(f"world {world} transport {opt.slabs.transport}: step dense {d:.3f} ms, sparse rows {s:.3f} ms")
opt.close()
if world > 1:
    dist.destroy_process_group()
