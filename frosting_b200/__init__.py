"""frosting_b200 -- B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.

A drop-in for the render path of Anttwo/Frosting: the `diff_gaussian_rasterization` Python surface
(GaussianRasterizationSettings / GaussianRasterizer) and the mesh occlusion-culling prepass, backed
by hand-written CUDA behind the C ABI of include/frosting_b200.h.  See DESIGN.md / INTEGRATION.md.
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,
                         forward_with_state)
from .frosting_attrs import frosting_attributes_fused
from .frosting_render import frosting_render, frosting_render_two_step
from .loss import l1_dssim_loss
from .optim import FrostingAdam, OptimizationParams
from .mesh import (MeshRasterizer, RasterizationSettings, Fragments, nvdiff_rasterization,
                   nvdiff_rasterization_with_pix_to_face, rasterize_mesh, gaussian_render_mask)

__all__ = [
    "GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "forward_with_state",
    "MeshRasterizer", "RasterizationSettings", "Fragments", "nvdiff_rasterization",
    "nvdiff_rasterization_with_pix_to_face", "rasterize_mesh", "gaussian_render_mask",
    "install_as_diff_gaussian_rasterization", "frosting_attributes_fused", "frosting_render", "frosting_render_two_step", "l1_dssim_loss", "FrostingAdam", "OptimizationParams",
]


def install_as_diff_gaussian_rasterization():
    """Make `import diff_gaussian_rasterization` resolve to this package (no edit of the caller's tree).

    Equivalent to putting frosting_b200/compat on sys.path; the reference's own backend switch is a
    source-level bool (frosting_scene/frosting_model.py:23), this is the import-site equivalent."""
    import sys
    from .compat import diff_gaussian_rasterization as shim
    sys.modules["diff_gaussian_rasterization"] = shim
    return shim
