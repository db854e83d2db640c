"""Host-side surface: same names, fields, argument checks and errors as the reference's Python module
(DGR/diff_gaussian_rasterization/__init__.py), scene generators, camera conventions, no CPU fallback."""
import math

import pytest
import torch

import frosting_b200 as fb
from frosting_b200 import scenes
from frosting_b200 import camera_batch as sharding


def test_settings_namedtuple_matches_reference_field_order():
    # DGR/diff_gaussian_rasterization/__init__.py:157-169
    assert fb.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    cam = scenes.make_camera(64, 48)
    rs = scenes.settings_for(cam, 2)
    assert rs.image_height == 48 and rs.image_width == 64 and rs.sh_degree == 2 and rs.debug is False
    r = fb.GaussianRasterizer(raster_settings=rs)
    assert isinstance(r, torch.nn.Module) and r.raster_settings is rs


def test_argument_validation_messages_match_reference():
    cam = scenes.make_camera(64, 48)
    r = fb.GaussianRasterizer(scenes.settings_for(cam, 0))
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(means3D=x, means2D=x, opacities=x[:, :1], scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(means3D=x, means2D=x, opacities=x[:, :1], shs=torch.zeros(4, 1, 3), colors_precomp=x, scales=x,
          rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, opacities=x[:, :1], colors_precomp=x)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, opacities=x[:, :1], colors_precomp=x, scales=x, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    cam = scenes.make_camera(64, 48)
    r = fb.GaussianRasterizer(scenes.settings_for(cam, 0))
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        r(means3D=x, means2D=x, opacities=x[:, :1], colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        r.markVisible(x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        fb.rasterize_mesh(x, torch.zeros(1, 3, dtype=torch.int32), torch.eye(4), 8, 8)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        r(means3D=torch.zeros(4, 2), means2D=x, opacities=x[:, :1], colors_precomp=x, scales=x,
          rotations=torch.zeros(4, 4))


def test_compat_alias_module():
    import sys
    shim = fb.install_as_diff_gaussian_rasterization()
    import diff_gaussian_rasterization as dgr
    assert dgr is shim and sys.modules["diff_gaussian_rasterization"] is shim
    assert dgr.GaussianRasterizer is fb.GaussianRasterizer
    assert dgr.GaussianRasterizationSettings is fb.GaussianRasterizationSettings


def test_camera_conventions():
    # frosting_utils/graphics_utils.py:65-85 and frosting_scene/cameras.py:203-212
    cam = scenes.make_camera(1920, 1080)
    assert math.isclose(cam.tanfovx, math.tan(math.radians(30)), rel_tol=1e-12)
    assert math.isclose(cam.tanfovy, cam.tanfovx * 1080 / 1920, rel_tol=1e-12)
    P = scenes.get_projection_matrix(0.01, 100.0, cam.FoVx, cam.FoVy)
    assert P[3, 2] == 1.0 and torch.allclose(P[0, 0], torch.tensor(1 / cam.tanfovx))
    assert torch.allclose(cam.full_proj_transform, cam.world_view_transform @ P.t())
    # a point on the optical axis projects to the image centre, w = depth
    p = torch.tensor([0.0, 0.0, 5.0, 1.0]) @ cam.full_proj_transform
    assert torch.allclose(p[:2], torch.zeros(2)) and math.isclose(float(p[3]), 5.0, rel_tol=1e-6)
    ring = scenes.ring_cameras(4, 64, 48, radius=10.0)
    for c in ring:
        centre_cam = torch.tensor([0.0, 0.0, 6.0, 1.0]) @ c.world_view_transform
        assert torch.allclose(centre_cam[:3], torch.tensor([0.0, 0.0, 10.0]), atol=1e-5)   # looks at the shell centre
        assert torch.allclose(c.camera_center, torch.linalg.inv(c.w2c)[:3, 3], atol=1e-5)


def test_scene_generators_are_deterministic_and_well_formed():
    cam = scenes.make_camera(320, 240)
    a, b = scenes.random_gaussians(5000, cam, 3), scenes.random_gaussians(5000, cam, 3)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["shs"].shape == (5000, 16, 3) and a["opacities"].shape == (5000, 1)
    assert torch.allclose(a["rotations"].norm(dim=1), torch.ones(5000), atol=1e-5)
    params, mesh = scenes.frosting_layer(4000, cam, 5, n_faces_target=3000)
    attrs = scenes.frosting_attributes(params, mesh)
    F = mesh["faces"].shape[0]
    assert int(mesh["cells"].max()) < F and mesh["cells"].shape == (4000,)
    # every Gaussian lies inside its shell (radius 3 +- thickness)
    d = (attrs["means3D"] - torch.tensor([0.0, 0.0, 6.0])).norm(dim=1)
    assert float(d.min()) > 3 - 0.05 and float(d.max()) < 3 + 0.021   # chord sag of the coarse test mesh
    # closed, outward-oriented base mesh
    f = mesh["faces"].long()
    e = torch.sort(torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1).values
    _, cnt = torch.unique(e, dim=0, return_counts=True)
    assert bool((cnt == 2).all())


def test_camera_block_partition():
    for world in (1, 2, 4, 8, 3):
        blocks = [sharding.camera_block(r, world, 64) for r in range(world)]
        flat = [c for b in blocks for c in b]
        assert flat == list(range(64))
        assert max(map(len, blocks)) - min(map(len, blocks)) <= 1
    with pytest.raises(ValueError):
        sharding.camera_block(2, 2, 8)
    assert float(sharding.reduce_loss(torch.tensor(3.0))) == 3.0   # no process group: identity


def test_capacity_hint_moves_rarely_and_in_coarse_steps():
    """`_grow_hint` (frosting_b200/rasterizer.py): every distinct capacity is a distinct allocation size, so the hint has
    hysteresis (untouched while >= 20 % headroom is left) and a coarse grid (1/8 of the leading power of two)."""
    from frosting_b200.rasterizer import _grow_hint, _HEADROOM
    h = _grow_hint(0, 1_000_000)
    assert h >= _HEADROOM * 1_000_000 and h % (1 << 17) == 0
    assert _grow_hint(h, 1_200_000) == h                      # creeping counts do not move it
    assert _grow_hint(h, int(0.8 * h)) == h
    h2 = _grow_hint(h, int(0.8 * h) + 1)                      # less than 20 % headroom left: grows, to 2x the count
    assert h2 > h and h2 >= 2 * (int(0.8 * h) + 1)
    assert _grow_hint(h2, 10) == h2                           # never shrinks
    sizes = {_grow_hint(0, n) for n in range(900_000, 1_100_000, 1000)}
    assert len(sizes) <= 3                                    # neighbouring counts share an allocation size
    assert _grow_hint(0, 0) >= 65536 and _grow_hint(0, -5) >= 65536


def test_grad_sink_row_bookkeeping():
    """GradSink (frosting_b200/optim.py): fetching a view marks the group written; sparse-row producers hand `radii` over
    only to sinks that declared row structure; `clear()` forgets both."""
    import torch
    from frosting_b200.optim import GradSink, slab_layout
    k = GradSink(a=torch.zeros(4, 3), b=torch.zeros(4))
    assert not k.written and k.row_radii is None
    _ = k["a"]
    assert k.written == {"a"} and k.get("b") is not None and k.written == {"a"}      # dict.get does not mark
    import pytest
    with pytest.raises(RuntimeError):
        k.rows_from(torch.ones(4, dtype=torch.int32), ["a"])                          # no row structure declared
    k.accepts_row_radii = True
    r = torch.tensor([1, 0, -1, 5], dtype=torch.int32)
    k.rows_from(r, ["a"])
    assert k.row_radii is r and k.sparse_names == {"a"}
    k.sink_once(["a"])
    with pytest.raises(RuntimeError):
        k.sink_once(["a", "b"])                                                       # a second overwrite before the step
    k.clear()
    assert not k.written and not k.sunk and k.row_radii is None and not k.sparse_names
    # the radii region sits behind the Adam range of the slab, 16-byte aligned
    starts, total = slab_layout([12, 4], world=2)
    assert total % 8 == 0 and starts == [0, 12, 16]
