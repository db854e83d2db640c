"""Row f3 on the GPU: fb200_adam_step against the numpy oracle and torch.optim.Adam, the grad_sink route of the fused
attribute backward, and (when the box has >= 2 GPUs) the peer-memory data-parallel step under NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = {"a": (1001, 6), "b": (1001, 1, 3), "c": (1001, 15, 3), "d": (1001, 1), "e": (333,), "f": (1001, 4)}
LRS = {"a": 0.005, "b": 0.0025, "c": 0.000125, "d": 0.05, "e": 0.005, "f": 0.001}


def _grads(t, rank, n):
    rng = np.random.default_rng(1000 * t + rank)
    g = rng.standard_normal(n).astype(np.float32) * (10.0 ** rng.integers(-6, 1, n)).astype(np.float32)
    g[rng.random(n) < 0.3] = 0.0
    return g


ROWS = 1001
ROW_NAMES = [n for n, sh in SHAPES.items() if sh[0] == ROWS]


def _radii(t, rank):
    """Which of the ROWS rows rank `rank` 'rendered' at step t (about a quarter, in runs like a real frame)."""
    rng = np.random.default_rng(77 * t + rank)
    r = (rng.random(ROWS) < 0.25).astype(np.int32) * rng.integers(1, 40, ROWS).astype(np.int32)
    r[100:300] = 0
    r[rng.random(ROWS) < 0.05] = -1      # radii are signed: anything <= 0 is "not rendered"
    return r


def _sparse_grads(t, rank, name, n):
    """Dense gradient of the step, what the slab holds (dead rows poisoned: they must never be read), and its radii."""
    g = _grads(t, rank, n)
    if name not in ROW_NAMES:
        return g, g
    radii = _radii(t, rank // 10 if rank >= 10 else 0)
    w = n // ROWS
    live = np.repeat(radii > 0, w)
    dense = np.where(live, g, 0.0).astype(np.float32)
    slab = np.where(live, g, np.nan).astype(np.float32)
    return dense, slab


def test_sparse_row_step_matches_the_dense_oracle():
    """fb200_adam_args.peer_row_radii: rows with radii <= 0 are zero rows that are NOT read -- the slab holds NaN there."""
    from frosting_b200 import optim
    from oracle import adam as adam_oracle
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    init = {n: rng.standard_normal(s).astype(np.float32) for n, s in SHAPES.items()}
    opt = optim.FrostingAdam({n: torch.from_numpy(x).to(dev) for n, x in init.items()}, LRS, rows=ROWS)
    assert opt.grads.accepts_row_radii and opt.slabs.row_width == [6, 3, 45, 1, 0, 4]
    ora = {n: (init[n].reshape(-1).copy(), np.zeros(init[n].size, np.float32), np.zeros(init[n].size, np.float32))
           for n in SHAPES}
    for t in range(1, 6):
        for k, n in enumerate(SHAPES):
            dense, slab = _sparse_grads(t, k, n, init[n].size)
            opt.grads[n].copy_(torch.from_numpy(slab).view(SHAPES[n]))
            ora[n] = adam_oracle.adam_step(ora[n][0], dense, ora[n][1], ora[n][2], LRS[n], t)
        opt.grads.rows_from(torch.from_numpy(_radii(t, 0)).to(dev), ROW_NAMES)
        opt.step()
        assert opt.grads.row_radii is None
    for n in SHAPES:
        got = opt.params[n].detach().cpu().numpy().reshape(-1)
        assert np.isfinite(got).all(), n
        np.testing.assert_allclose(got, ora[n][0], rtol=3e-6, atol=5e-7)


def test_fused_adam_matches_oracle_and_torch():
    from frosting_b200 import optim
    from oracle import adam as adam_oracle
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    init = {n: rng.standard_normal(s).astype(np.float32) for n, s in SHAPES.items()}
    opt = optim.FrostingAdam({n: torch.from_numpy(x).to(dev) for n, x in init.items()}, LRS)
    ref_p = {n: torch.from_numpy(x.copy()).to(dev).requires_grad_(True) for n, x in init.items()}
    ref = torch.optim.Adam([{"params": [ref_p[n]], "lr": LRS[n]} for n in SHAPES], lr=0.0, eps=1e-15)
    ora = {n: (init[n].reshape(-1).copy(), np.zeros(init[n].size, np.float32), np.zeros(init[n].size, np.float32))
           for n in SHAPES}
    before = __import__("frosting_b200")._lib.kernel_launches()
    for t in range(1, 6):
        for k, n in enumerate(SHAPES):
            g = _grads(t, k, init[n].size)
            opt.grads[n].copy_(torch.from_numpy(g).view(SHAPES[n]))
            ref_p[n].grad = torch.from_numpy(g.copy()).view(SHAPES[n]).to(dev)
            ora[n] = adam_oracle.adam_step(ora[n][0], g, ora[n][1], ora[n][2], LRS[n], t)
        opt.step()
        ref.step()
    assert __import__("frosting_b200")._lib.kernel_launches() - before == 5       # one kernel per step, all groups
    for n in SHAPES:
        got = opt.params[n].detach().cpu().numpy().reshape(-1)
        np.testing.assert_allclose(got, ora[n][0], rtol=3e-6, atol=5e-7)
        np.testing.assert_allclose(got, ref_p[n].detach().cpu().numpy().reshape(-1), rtol=3e-6, atol=5e-7)
    # the slab padding between groups stays zero
    s = opt.slabs
    mask = torch.ones(s.total, dtype=torch.bool, device=dev)
    for n, st in zip(s.names, s.starts):
        mask[st:st + init[n].size] = False
    assert float(s.param_slab[mask].abs().max() if mask.any() else 0.0) == 0.0


def test_adam_argument_checks():
    import ctypes as C
    from frosting_b200 import _lib
    a = _lib.AdamArgs()
    a.world, a.rank, a.n_groups = 1, 0, 1
    a.group_start[0], a.group_start[1] = 0, 6          # not a multiple of 4
    with pytest.raises(_lib.Fb200Error):
        _lib.check(_lib.lib().fb200_adam_step(C.byref(a), None))
    a.group_start[1] = 8
    a.shard_lo, a.shard_hi = 0, 8                      # missing buffers
    a.bias_correction1 = a.bias_correction2_sqrt = 1.0
    with pytest.raises(_lib.Fb200Error):
        _lib.check(_lib.lib().fb200_adam_step(C.byref(a), None))
    a.world = 9
    with pytest.raises(_lib.Fb200Error):
        _lib.check(_lib.lib().fb200_adam_step(C.byref(a), None))


def test_grad_sink_receives_the_autograd_gradients():
    import frosting_b200 as fb
    from frosting_b200 import scenes, optim
    dev = torch.device("cuda:0")
    P = 20_000
    cam = scenes.make_camera(200, 120, device=dev)
    params, mesh = scenes.frosting_layer(P, cam, 5, n_faces_target=4000, device=dev)
    mask = (torch.rand(P, device=dev) < 0.7)
    w = mask.float()
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    out = fb.frosting_attributes_fused(leaf, mesh, mask=mask)
    cot = {k: torch.randn_like(v) for k, v in out.items()}
    sum(((out[k] * cot[k]).reshape(P, -1).sum(1) * w).sum() for k in out).backward()
    opt = optim.FrostingAdam.for_frosting(params)
    out2 = fb.frosting_attributes_fused(opt.params, mesh, mask=mask, grad_sink=opt.grads)
    sum(((out2[k] * cot[k]).reshape(P, -1).sum(1) * w).sum() for k in out2).backward()
    for k in ("bary_logits", "opacity_logits", "log_scales", "quats", "sh_dc", "sh_rest"):
        assert opt.params[k].grad is None
        assert torch.equal(opt.grads[k].reshape(-1), leaf[k].grad.reshape(-1)), k
    assert [g["name"] for g in opt.param_groups] == ["bary_coords", "sh_coordinates_dc", "sh_coordinates_rest",
                                                     "opacities", "scales", "quaternions"]
    opt.step()
    assert opt.update_learning_rate(15_000) == pytest.approx(np.sqrt(0.005 * 0.00005))


def test_frosting_render_sparse_sink_step_equals_the_dense_step():
    """frosting_render -> FrostingAdam without a dense gradient anywhere: the backward writes the rendered rows of the
    optimizer's slab (poisoned with NaN beforehand), hands `radii` over, and the step equals the step on the dense copy."""
    import frosting_b200 as fb
    from frosting_b200 import scenes, optim
    dev = torch.device("cuda:0")
    P, W, H = 30_000, 240, 160
    cam = scenes.make_camera(W, H, device=dev)
    params, mesh = scenes.frosting_layer(P, cam, 5, n_faces_target=5000, device=dev, view_distance=4.5)
    rs = scenes.settings_for(cam, 3, device=dev)
    _, fv, _ = fb.rasterize_mesh(mesh["verts"], mesh["faces"], cam.full_proj_transform, H, W, mark_last_on_bg=True)
    cot = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev)
    a, b = optim.FrostingAdam.for_frosting(params), optim.FrostingAdam.for_frosting(params)
    for step in range(2):
        a.slabs.grad_slab.fill_(float("nan"))
        color, radii = fb.frosting_render(a.params, mesh, rs, face_visible=fv, grad_sink=a.grads)
        (color * cot).sum().backward()
        assert a.grads.row_radii is not None and set(a.grads.sparse_names) == set(a.slabs.names)
        live = radii > 0
        assert 0 < int(live.sum()) < P
        for n in a.slabs.names:
            g = a.slabs.grads.get(n)                       # dict.get: does not touch the book-keeping
            assert bool(torch.isfinite(g[live]).all()), n
            assert bool(torch.isnan(g[~live]).all()), n    # unrendered rows really were not written
            b.grads[n].copy_(torch.where(live.view(-1, *([1] * (g.dim() - 1))), g, torch.zeros_like(g)))
        a.step(); b.step()
        assert torch.equal(a.slabs.param_slab[:a.slabs.total].view(torch.int32), b.slabs.param_slab[:b.slabs.total].view(torch.int32))
        assert torch.equal(a.exp_avg, b.exp_avg) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
        assert bool(torch.isfinite(a.slabs.param_slab).all())


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _peer_worker(rank, world, port, q, multicast, sparse=False):
    import torch.distributed as dist
    from frosting_b200 import optim
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["FB200_MULTICAST"] = "1" if multicast else "0"     # force / forbid the in-switch path at world 2
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rng = np.random.default_rng(3)
    init = {n: rng.standard_normal(s).astype(np.float32) for n, s in SHAPES.items()}
    opt = optim.FrostingAdam({n: torch.from_numpy(x).to(dev) for n, x in init.items()}, LRS, rows=ROWS if sparse else None)
    loss_sum = None
    for t in range(1, 4):
        for k, n in enumerate(SHAPES):
            g = _sparse_grads(t, 10 * rank + k, n, init[n].size)[1] if sparse else _grads(t, 10 * rank + k, init[n].size)
            opt.grads[n].copy_(torch.from_numpy(g).view(SHAPES[n]))
        if sparse:
            opt.grads.rows_from(torch.from_numpy(_radii(t, rank)).to(dev), ROW_NAMES)
        loss = torch.tensor([float(rank + 1)], device=dev)
        opt.step(loss=loss)
        loss_sum = float(loss)
    torch.cuda.synchronize(dev)
    q.put((rank, {n: opt.params[n].detach().cpu().numpy().reshape(-1) for n in SHAPES}, loss_sum, opt.slabs.transport))
    opt.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="peer-memory step needs 2 GPUs")
@pytest.mark.parametrize("sparse", [False, True], ids=["dense", "sparse-rows"])
@pytest.mark.parametrize("multicast", [False, True], ids=["peer-pointers", "nvswitch-multicast"])
def test_peer_memory_dp_adam_world2(multicast, sparse):
    import torch.multiprocessing as mp
    from oracle import adam as adam_oracle
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q, multicast, sparse)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, params, loss_sum, transport = q.get(timeout=300)
        res[r] = params
        assert loss_sum == 3.0
        print("transport:", transport)
        assert ("multicast" in transport) <= multicast            # never the in-switch path when forbidden
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    init = {n: rng.standard_normal(s).astype(np.float32) for n, s in SHAPES.items()}
    for k, n in enumerate(SHAPES):
        assert np.array_equal(res[0][n], res[1][n]), n          # replicas bit-identical
        p, m, v = init[n].reshape(-1).copy(), np.zeros(init[n].size, np.float32), np.zeros(init[n].size, np.float32)
        for t in range(1, 4):
            gs = [_sparse_grads(t, 10 * r + k, n, init[n].size)[0] if sparse else _grads(t, 10 * r + k, init[n].size)
                  for r in range(world)]
            p, m, v = adam_oracle.dp_step(p, gs, m, v, LRS[n], t, 0.5)
        np.testing.assert_allclose(res[0][n], p, rtol=3e-6, atol=5e-7)


def test_state_dict_round_trips_through_torch_adam():
    """the reference checkpoints `optimizer.state_dict()` = torch.optim.Adam's (frosting_optimizer.py:139,
    refine.py:543-551).  Ours has the same format: it loads into a torch Adam over the same groups and back, and both
    continue identically."""
    from frosting_b200 import optim
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    init = {n: rng.standard_normal(s).astype(np.float32) for n, s in SHAPES.items()}
    opt = optim.FrostingAdam({n: torch.from_numpy(x).to(dev) for n, x in init.items()}, LRS)
    assert opt.state_dict()["state"] == {}
    for t in range(1, 4):
        for k, n in enumerate(SHAPES):
            opt.grads[n].copy_(torch.from_numpy(_grads(t, k, init[n].size)).view(SHAPES[n]))
        opt.step()
    sd = opt.state_dict()
    assert sorted(sd) == ["param_groups", "state"] and len(sd["param_groups"]) == len(SHAPES)
    # -> torch.optim.Adam over the same values
    ref_p = {n: opt.params[n].detach().clone().requires_grad_(True) for n in SHAPES}
    ref = torch.optim.Adam([{"params": [ref_p[n]], "lr": LRS[n]} for n in SHAPES], lr=0.0, eps=1e-15)
    ref.load_state_dict(sd)
    # -> a fresh FrostingAdam from what torch saves
    opt2 = optim.FrostingAdam({n: opt.params[n].detach().clone() for n in SHAPES}, {n: 0.0 for n in SHAPES})
    opt2.load_state_dict(ref.state_dict())
    assert opt2.current_iteration == 3 and [g["lr"] for g in opt2.param_groups] == [LRS[n] for n in SHAPES]
    for k, n in enumerate(SHAPES):
        g = _grads(9, k, init[n].size)
        for o in (opt, opt2):
            o.grads[n].copy_(torch.from_numpy(g).view(SHAPES[n]))
        ref_p[n].grad = torch.from_numpy(g.copy()).view(SHAPES[n]).to(dev)
    opt.step(); opt2.step(); ref.step()
    for n in SHAPES:
        a = opt.params[n].detach().cpu().numpy().reshape(-1)
        np.testing.assert_array_equal(a, opt2.params[n].detach().cpu().numpy().reshape(-1))
        np.testing.assert_allclose(a, ref_p[n].detach().cpu().numpy().reshape(-1), rtol=3e-6, atol=5e-7)


def test_groups_without_a_gradient_are_skipped_and_double_sink_raises():
    """torch.optim.Adam skips parameters whose .grad is None; a stale slab must never be applied twice."""
    import frosting_b200 as fb
    from frosting_b200 import optim, scenes
    dev = torch.device("cuda:0")
    init = {n: torch.randn(s, generator=torch.Generator().manual_seed(1)).to(dev) for n, s in SHAPES.items()}
    opt = optim.FrostingAdam(init, LRS)
    for n in SHAPES:
        opt.grads[n].fill_(0.5)
    opt.step()
    after1 = {n: opt.params[n].detach().clone() for n in SHAPES}
    # second step: only "a" and "d" get a gradient (through autograd's .grad, collected by step())
    opt.zero_grad()
    (opt.params["a"].sum() * 2.0 + opt.params["d"].sum()).backward()
    opt.step()
    for n in SHAPES:
        same = torch.equal(opt.params[n].detach(), after1[n])
        assert same == (n not in ("a", "d")), n
    # two fused backwards into the sink between steps
    P = 2_000
    cam = scenes.make_camera(64, 48, device=dev)
    params, mesh = scenes.frosting_layer(P, cam, 5, n_faces_target=500, device=dev)
    fo = optim.FrostingAdam.for_frosting(params)
    for i in range(2):
        out = fb.frosting_attributes_fused(fo.params, mesh, grad_sink=fo.grads)
        loss = sum(v.sum() for v in out.values())
        if i == 0:
            loss.backward()
        else:
            with pytest.raises(RuntimeError, match="written twice"):
                loss.backward()
    fo.step()
    out = fb.frosting_attributes_fused(fo.params, mesh, grad_sink=fo.grads)
    sum(v.sum() for v in out.values()).backward()        # after a step the sink accepts the next backward
