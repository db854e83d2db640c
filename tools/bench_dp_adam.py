"""Kernel-only timing of the peer-memory reduce + Adam + publish step (row f3).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dp_adam.py [P]
Prints, per world size: ms per step (max over ranks, CUDA events around 20 launches, no rendezvous inside the timed
region) and the NVLink payload rate per GPU and direction."""
import os
import sys
import ctypes as C

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frosting_b200 import optim, _lib   # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shapes = {"bary_logits": (P, 6), "sh_dc": (P, 1, 3), "sh_rest": (P, 15, 3), "opacity_logits": (P,),
              "log_scales": (P, 3), "quats": (P, 4)}
    g = torch.Generator().manual_seed(1)
    params = {k: torch.randn(s, generator=g).to(dev) for k, s in shapes.items()}
    opt = optim.FrostingAdam.for_frosting(params)
    for k in opt.grads:
        opt.grads[k].normal_()
    n = opt.slabs.total
    stream = torch.cuda.current_stream(dev)
    L = _lib.lib()

    def launch():
        opt.current_iteration += 1
        a = opt._args()
        _lib.check(L.fb200_adam_step(C.byref(a), C.c_void_p(stream.cuda_stream)))

    for _ in range(3):
        launch()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record()
    for _ in range(iters):
        launch()
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / iters
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    shard = (opt.hi - opt.lo) * 4
    remote = shard * (world - 1)                 # bytes in (gradients) = bytes out (parameters), per GPU
    hbm = shard * (1 + 3 + 3) + 2 * remote       # own grad + m,v,p reads + m,v,p writes + what the peers read / write here
    if rank == 0:
        print(f"world {world} [{opt.slabs.transport}]: {n / 1e6:.1f} M parameters, {ms:.3f} ms per step; NVLink {remote / 1e6:.0f} MB each way per GPU "
              f"-> {remote / ms / 1e6:.0f} GB/s per direction; local HBM traffic {hbm / 1e6:.0f} MB -> {hbm / ms / 1e6:.0f} GB/s")
    if world > 1:
        dist.barrier()
    opt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
