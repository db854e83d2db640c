"""Loader + thin driver for the UNMODIFIED reference rasterizer built by oracle/build_ref.py.

Test infrastructure only (see oracle/__init__.py).  Calls the reference's own pybind entry points
(DGR/ext.cpp:15-19) with the argument order of DGR/diff_gaussian_rasterization/__init__.py:59-79 and
:110-130, and exposes the reference's internal buffers by re-deriving the offsets of obtain()
(DGR/cuda_rasterizer/rasterizer_impl.h:22-28, rasterizer_impl.cu:155-194) -- the oracle itself is not
edited.
"""
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "ref_dgr_C.so")
_mod = None


def available() -> bool:
    return os.path.exists(SO)


def module():
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{SO} missing: run `python oracle/build_ref.py` where /root/reference exists")
        spec = importlib.util.spec_from_file_location("ref_dgr_C", SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def _e(t, device):
    return torch.empty(0, device="cpu") if t is None else t


def forward(rs, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None):
    """rs: any object with the 12 GaussianRasterizationSettings fields.  Returns dict."""
    C = module()
    e = torch.Tensor([])
    args = (rs.bg, means3D, e if colors_precomp is None else colors_precomp, opacities,
            e if scales is None else scales, e if rotations is None else rotations, rs.scale_modifier,
            e if cov3D_precomp is None else cov3D_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            rs.image_height, rs.image_width, e if shs is None else shs, rs.sh_degree, rs.campos, rs.prefiltered,
            rs.debug)
    num_rendered, color, radii, geom, binning, img = C.rasterize_gaussians(*args)
    return dict(num_rendered=num_rendered, color=color, radii=radii, geom=geom, binning=binning, img=img,
                args=args)


def backward(rs, fwd, means3D, grad_out_color, shs=None, colors_precomp=None, scales=None, rotations=None,
             cov3D_precomp=None):
    C = module()
    e = torch.Tensor([])
    args = (rs.bg, means3D, fwd["radii"], e if colors_precomp is None else colors_precomp,
            e if scales is None else scales, e if rotations is None else rotations, rs.scale_modifier,
            e if cov3D_precomp is None else cov3D_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            grad_out_color, e if shs is None else shs, rs.sh_degree, rs.campos, fwd["geom"], fwd["num_rendered"],
            fwd["binning"], fwd["img"], rs.debug)
    (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot) = C.rasterize_gaussians_backward(*args)
    return dict(means2D=g_means2D, colors=g_colors, opacities=g_opac, means3D=g_means3D, cov3D=g_cov3D, sh=g_sh,
                scales=g_scales, rotations=g_rot)


class RefRasterize(torch.autograd.Function):
    """fwd+bwd through the reference's entry points, for timing the reference arm."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, opacities, scales, rotations, rs):
        fwd = forward(rs, means3D, opacities, shs=shs, scales=scales, rotations=rotations)
        ctx.rs, ctx.fwd = rs, fwd
        ctx.save_for_backward(means3D, shs, scales, rotations)
        return fwd["color"], fwd["radii"]

    @staticmethod
    def backward(ctx, g, _):
        means3D, shs, scales, rotations = ctx.saved_tensors
        b = backward(ctx.rs, ctx.fwd, means3D, g, shs=shs, scales=scales, rotations=rotations)
        return b["means3D"], b["means2D"], b["sh"], b["opacities"], b["scales"], b["rotations"], None


def _al(x, a=128):
    return (x + a - 1) // a * a


def geom_views(geom: torch.Tensor, P: int):
    """Arrays of GeometryState in obtain() order (rasterizer_impl.cu:155-170)."""
    base = geom.data_ptr()
    assert base % 128 == 0
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = geom[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes
    take("depths", 4 * P, torch.float32, (P,))
    take("clamped", 3 * P, torch.uint8, (P, 3))
    take("internal_radii", 4 * P, torch.int32, (P,))
    take("means2D", 8 * P, torch.float32, (P, 2))
    take("cov3D", 24 * P, torch.float32, (P, 6))
    take("conic_opacity", 16 * P, torch.float32, (P, 4))
    take("rgb", 12 * P, torch.float32, (P, 3))
    take("tiles_touched", 4 * P, torch.int32, (P,))
    # point_offsets sits after the CUB scan temp; locate it from the end (required() adds 128)
    po = geom.numel() - 128 - 4 * P
    assert po % 128 == 0 and po >= off, (po, off)
    out["point_offsets"] = geom[po:po + 4 * P].view(torch.int32)
    return out


def binning_views(binning: torch.Tensor, R: int):
    off = 0
    out = {}

    def take(name, nbytes, dtype):
        nonlocal off
        off = _al(off)
        out[name] = binning[off:off + nbytes].view(dtype)
        off += nbytes
    take("point_list", 4 * R, torch.int32)
    take("point_list_unsorted", 4 * R, torch.int32)
    take("point_list_keys", 8 * R, torch.int64)
    take("point_list_keys_unsorted", 8 * R, torch.int64)
    return out


def img_views(img: torch.Tensor, H: int, W: int):
    N = H * W
    T = ((W + 15) // 16) * ((H + 15) // 16)
    off = 0
    out = {}

    def take(name, nbytes, dtype):
        nonlocal off
        off = _al(off)
        out[name] = img[off:off + nbytes].view(dtype)
        off += nbytes
    take("accum_alpha", 4 * N, torch.float32)
    take("n_contrib", 4 * N, torch.int32)
    take("ranges", 8 * N, torch.int32)
    out["accum_alpha"] = out["accum_alpha"].view(H, W)
    out["n_contrib"] = out["n_contrib"].view(H, W)
    out["ranges"] = out["ranges"].view(N, 2)[:T]
    return out
