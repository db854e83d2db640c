"""Row f3 on CPU: the Adam oracle against torch.optim.Adam's golden vectors, the slab layout / shard logic of
frosting_b200/optim.py, and the data-parallel statement (reduce own shard -> update -> publish) under gloo, world 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frosting_b200 import optim
from oracle import adam as adam_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "adam.npz")


def test_adam_oracle_matches_torch_golden():
    z = np.load(GOLD)
    for n in z["names"]:
        p = z[f"p0_{n}"]
        m, v = np.zeros_like(p), np.zeros_like(p)
        for t in range(int(z["steps"])):
            p, m, v = adam_oracle.adam_step(p, z[f"g{t}_{n}"], m, v, float(z[f"lr_{n}"]), t + 1)
            ref = z[f"p{t + 1}_{n}"]
            # same formulas in fp32; torch contracts some mul+add pairs differently -> a few ulp
            np.testing.assert_allclose(p, ref, rtol=2e-6, atol=1e-7)


def test_slab_layout_and_shards():
    starts, total = optim.slab_layout([6 * 37, 3 * 37, 45 * 37, 37, 3 * 37, 4 * 37], world=8)
    assert all(s % 4 == 0 for s in starts) and total % 32 == 0 and total >= starts[-1]
    for a, b, n in zip(starts, starts[1:], [222, 111, 1665, 37, 111, 148]):
        assert b - a >= n and b - a < n + 4
    seen = []
    for r in range(8):
        lo, hi = optim.shard_range(total, r, 8)
        assert lo % 4 == 0 and hi % 4 == 0
        seen += list(range(lo, hi, 4))
    assert seen == list(range(0, total, 4))
    assert optim.slab_layout([], 1) == ([0], 0)
    with pytest.raises(ValueError):
        optim.shard_range(10, 0, 2)
    with pytest.raises(ValueError):
        optim.shard_range(16, 2, 2)


def test_lr_schedule_matches_reference_formula():
    f = optim.get_expon_lr_func(0.005, 0.00005, lr_delay_mult=0.01, max_steps=30_000)
    assert f(0) == pytest.approx(0.005) and f(30_000) == pytest.approx(0.00005) and f(60_000) == pytest.approx(0.00005)
    assert f(15_000) == pytest.approx(np.sqrt(0.005 * 0.00005))
    g = optim.get_expon_lr_func(1e-2, 1e-4, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=1000)
    assert g(0) == pytest.approx(0.1 * 1e-2) and g(-1) == 0.0
    assert optim.get_expon_lr_func(0.0, 0.0)(5) == 0.0


def test_frosting_adam_refuses_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        optim.FrostingAdam({"w": torch.zeros(8)}, {"w": 0.1})


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    sizes = [30, 7, 13]
    lrs = [0.05, 0.001, 0.005]
    starts, total = optim.slab_layout(sizes, world)
    lo, hi = optim.shard_range(total, rank, world)
    p = rng.standard_normal(total).astype(np.float32)           # same on both ranks
    m, v = np.zeros(hi - lo, np.float32), np.zeros(hi - lo, np.float32)
    lr_of = np.zeros(total, np.float32)
    for k, s in enumerate(starts[:-1]):
        lr_of[s:] = lrs[k]
    for t in range(1, 4):
        g_all = [np.random.default_rng(100 * t + r).standard_normal(total).astype(np.float32) for r in range(world)]
        # what the kernel does, with gloo standing in for the NVLink loads / stores: gather the peers' shard of the
        # gradients, update the shard element-wise (per-element lr), publish the shard of the parameters
        mine = torch.from_numpy(g_all[rank].copy())
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        new = p[lo:hi].copy()
        for k in range(len(sizes)):
            a, b = max(starts[k], lo), min(starts[k + 1] if k + 1 < len(sizes) else total, hi)
            if a < b:
                new[a - lo:b - lo], m[a - lo:b - lo], v[a - lo:b - lo] = adam_oracle.dp_step(
                    p[a:b], [x.numpy()[a:b] for x in gathered], m[a - lo:b - lo], v[a - lo:b - lo], lrs[k], t, 1.0 / world)
        shards = [torch.empty(hi - lo) for _ in range(world)]
        dist.all_gather(shards, torch.from_numpy(new))
        p = torch.cat(shards).numpy()
    q.put((rank, p))
    dist.destroy_process_group()


def test_sharded_dp_adam_equals_single_process_adam_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert np.array_equal(res[0], res[1])                          # replicas stay bit-identical
    # single-process torch.optim.Adam on the averaged gradients
    sizes, lrs = [30, 7, 13], [0.05, 0.001, 0.005]
    starts, total = optim.slab_layout(sizes, world)
    p0 = np.random.default_rng(7).standard_normal(total).astype(np.float32)
    groups = [torch.from_numpy(p0[s:s + n].copy()).requires_grad_(True) for s, n in zip(starts, sizes)]
    opt = torch.optim.Adam([{"params": [g], "lr": lr} for g, lr in zip(groups, lrs)], lr=0.0, eps=1e-15)
    for t in range(1, 4):
        g_all = [np.random.default_rng(100 * t + r).standard_normal(total).astype(np.float32) for r in range(world)]
        avg = (g_all[0] + g_all[1]) * np.float32(0.5)
        for g, s, n in zip(groups, starts, sizes):
            g.grad = torch.from_numpy(avg[s:s + n].copy())
        opt.step()
    for g, s, n in zip(groups, starts, sizes):
        np.testing.assert_allclose(res[0][s:s + n], g.detach().numpy(), rtol=3e-6, atol=1e-7)
