"""Alias module: `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(frosting_scene/frosting_model.py:29, sugar_model.py:10, gaussian_renderer/__init__.py:14) keeps working
when this directory's parent is on sys.path or after frosting_b200.install_as_diff_gaussian_rasterization()."""
from frosting_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                      rasterize_gaussians, _RasterizeGaussians, cpu_deep_copy_tuple)
