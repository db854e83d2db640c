"""Camera-batch sharding across the GPUs of one box.

The render path shards by camera: every rank holds the full set of Gaussians and renders a contiguous
block of the camera batch; the only exchange is the all-reduce of the scalar loss.  The reference has no
multi-GPU code at all (single `--gpu` index, frosting_trainers/refine.py:26,249), so this is new surface.
"""
import torch
import torch.distributed as dist


def camera_block(rank: int, world: int, n_cameras: int):
    """Contiguous block of camera indices owned by `rank` (blocks differ by at most one camera)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_cameras, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def reduce_loss(local_loss: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank scalar losses on every rank (NCCL over NVLink on GPUs, gloo on CPU)."""
    out = local_loss.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
