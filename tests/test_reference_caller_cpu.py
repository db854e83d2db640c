"""Drop-in through the REFERENCE'S OWN caller (task 10).

`gaussian_splatting/gaussian_renderer/__init__.py:18-100` (the reference's `render`) is imported from /root/reference
-- unmodified, where it lies -- after `install_as_diff_gaussian_rasterization()`, so its
`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer` binds to this package.
It is then called with a stand-in GaussianModel / camera / pipe.  What this proves: the reference's keyword usage, argument
order, None-handling and the (image, radii) return bind to our surface without touching the caller.  The native call at
the bottom of our shim is replaced by a recorder here (this container has no GPU, and /root/reference does not exist on
the GPU box, so the two cannot meet); the numerics behind exactly this entry (`GaussianRasterizer.forward`) are what
tests/test_parity_gpu.py and tests/test_bench_configs_gpu.py compare with the compiled reference.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

import frosting_b200 as fb
from frosting_b200 import rasterizer as fbr

REF = "/root/reference/gaussian_splatting/gaussian_renderer/__init__.py"

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="/root/reference not present (GPU box)")


def _load_reference_renderer(monkeypatch):
    shim = fb.install_as_diff_gaussian_rasterization()
    monkeypatch.setitem(sys.modules, "diff_gaussian_rasterization", shim)
    # the caller's other imports: only names are needed (scene.gaussian_model.GaussianModel as an annotation,
    # utils.sh_utils.eval_sh on the convert_SHs_python branch)
    scene = types.ModuleType("scene")
    gm = types.ModuleType("scene.gaussian_model")
    gm.GaussianModel = object
    utils = types.ModuleType("utils")
    sh = types.ModuleType("utils.sh_utils")
    sh.eval_sh = lambda deg, shs, dirs: shs[..., 0]
    for name, mod in (("scene", scene), ("scene.gaussian_model", gm), ("utils", utils), ("utils.sh_utils", sh)):
        monkeypatch.setitem(sys.modules, name, mod)
    spec = importlib.util.spec_from_file_location("ref_gaussian_renderer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _PC:
    """The attributes of GaussianModel that render() reads (scene/gaussian_model.py:24-406)."""

    def __init__(self, P):
        g = torch.Generator().manual_seed(0)
        self.get_xyz = torch.randn(P, 3, generator=g)
        self.get_opacity = torch.rand(P, 1, generator=g)
        self.get_scaling = torch.rand(P, 3, generator=g)
        self.get_rotation = torch.randn(P, 4, generator=g)
        self.get_features = torch.randn(P, 16, 3, generator=g)
        self.active_sh_degree = 3
        self.max_sh_degree = 3

    def get_covariance(self, scaling_modifier=1):
        return torch.rand(self.get_xyz.shape[0], 6)


def test_reference_render_binds_to_our_surface(monkeypatch):
    mod = _load_reference_renderer(monkeypatch)
    assert mod.GaussianRasterizer is fb.GaussianRasterizer
    assert mod.GaussianRasterizationSettings is fb.GaussianRasterizationSettings

    # the caller creates its means2D sink with device="cuda" (:27); no CUDA in this container
    real_zeros_like = torch.zeros_like
    monkeypatch.setattr(torch, "zeros_like", lambda t, **kw: real_zeros_like(t, **{k: v for k, v in kw.items() if k != "device"}))

    seen = {}

    def recorder(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                 *extensions):
        seen.update(means3D=means3D, means2D=means2D, sh=sh, colors_precomp=colors_precomp, opacities=opacities,
                    scales=scales, rotations=rotations, cov3D=cov3Ds_precomp, rs=raster_settings, ext=extensions)
        H, W = raster_settings.image_height, raster_settings.image_width
        return torch.zeros(3, H, W), torch.ones(means3D.shape[0], dtype=torch.int32)

    monkeypatch.setattr(fbr, "rasterize_gaussians", recorder)
    from frosting_b200 import scenes
    cam = scenes.make_camera(64, 48)
    pc = _PC(100)
    bg = torch.tensor([0.0, 0.5, 1.0])

    # SH + scale/rotation path (the default pipeline flags, arguments/__init__.py:64-69)
    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    out = mod.render(cam, pc, pipe, bg, scaling_modifier=0.7)
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii"}
    assert out["render"].shape == (3, 48, 64) and bool(out["visibility_filter"].all())
    rs = seen["rs"]
    assert isinstance(rs, fb.GaussianRasterizationSettings)
    assert (rs.image_height, rs.image_width, rs.sh_degree, rs.prefiltered, rs.debug) == (48, 64, 3, False, False)
    assert rs.scale_modifier == 0.7 and rs.bg is bg and rs.viewmatrix is cam.world_view_transform
    assert rs.projmatrix is cam.full_proj_transform and rs.campos is cam.camera_center
    assert abs(rs.tanfovx - cam.tanfovx) < 1e-6 and abs(rs.tanfovy - cam.tanfovy) < 1e-6
    assert seen["means3D"] is pc.get_xyz and seen["sh"] is pc.get_features and seen["opacities"] is pc.get_opacity
    assert seen["scales"] is pc.get_scaling and seen["rotations"] is pc.get_rotation
    # the reference's None -> empty-tensor convention (__init__.py:197-207) is applied by OUR forward
    assert seen["colors_precomp"].numel() == 0 and seen["cov3D"].numel() == 0
    assert seen["means2D"] is out["viewspace_points"] and seen["means2D"].requires_grad
    assert not any(seen["ext"]), "a reference caller must never trigger an extension"

    # precomputed colour + precomputed covariance path
    pipe = types.SimpleNamespace(debug=True, compute_cov3D_python=True, convert_SHs_python=False)
    col = torch.rand(100, 3)
    mod.render(cam, pc, pipe, bg, override_color=col)
    assert seen["colors_precomp"] is col and seen["sh"].numel() == 0
    assert seen["cov3D"].shape == (100, 6) and seen["scales"].numel() == 0 and seen["rotations"].numel() == 0
    assert seen["rs"].debug is True


def test_invalid_combinations_raise_like_the_reference(monkeypatch):
    r = fb.GaussianRasterizer(None)
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=m[:, :1], scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=m, means2D=m, opacities=m[:, :1], shs=torch.zeros(4, 16, 3))
