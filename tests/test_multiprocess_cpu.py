"""N>1 host logic on CPU: gloo, world_size 2 -- camera-batch partition + scalar loss all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frosting_b200 import camera_batch as sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cams = sharding.camera_block(rank, world, 16)
    # stand-in for the per-camera render loss: a deterministic function of the camera index
    local = torch.tensor(float(sum((c + 1) ** 2 for c in cams)))
    total = sharding.reduce_loss(local)
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)             # the max-over-ranks timing reduction of bench.py
    # the per-frame loss reduction of CameraBatch.run: the frames' losses are buffered and go out `every` frames per
    # asynchronous all-reduce; whatever the batch size, frame k comes back as the sum of every rank's frame k, in order
    red = sharding.LossReducer()
    for c in cams:
        red.add(torch.tensor(float((c + 1) ** 2)))
    per_frame = red.collect().tolist()
    assert red.collect().numel() == 0
    for every in (1, 3):                                 # a partial last batch is flushed by collect()
        red = sharding.LossReducer(every=every)
        for c in cams:
            red.add(torch.tensor(float((c + 1) ** 2)))
        assert len(red.pending) == len(cams) // every
        assert red.collect().tolist() == per_frame
    q.put((rank, cams, float(local), float(total), float(t), per_frame))
    dist.destroy_process_group()


def test_camera_sharding_and_loss_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cams = [c for r in res for c in r[1]]
    assert cams == list(range(16))
    expect = float(sum((c + 1) ** 2 for c in range(16)))
    assert all(r[3] == expect for r in res) and res[0][2] + res[1][2] == expect
    assert all(r[4] == 2.0 for r in res)
    # frame k of rank 0 is reduced with frame k of rank 1 (cameras k and 8 + k)
    expect_frames = [float((k + 1) ** 2 + (8 + k + 1) ** 2) for k in range(8)]
    assert all(r[5] == expect_frames for r in res)
