"""Frosting's whole render call as ONE differentiable op.

`frosting_render(params, mesh, raster_settings, face_visible)` is what `Frosting.render_image_gaussian_rasterizer` does
between its learnable parameters and the image (frosting_scene/frosting_model.py:1470-1519 attribute properties :713-799,
:1564-1586 occlusion mask and boolean gathers, :1624 means2D sink, :1649-1657 rasterizer call), with the pieces that exist
only to glue torch ops together removed:

  * occlusion culling is a lookup `face_visible[cell[i]]` inside the attribute kernel and inside preprocess -- no
    `_index_mask`, no `render_mask`, no boolean gathers of five attribute tensors and no scatter in their backward;
  * the attribute tensors are produced by one kernel (csrc/frosting_attr.cu) for the visible Gaussians only;
  * the rasterizer's backward runs with the SPARSE-ROW contract (fb200_grads.sparse_rows): it does not write the zero
    rows of the ~75-90 % of Gaussians that were not rendered (round 1: 74 % of its DRAM traffic), because the next
    kernel -- the attribute backward, given `radii` -- never reads them; that kernel writes every parameter gradient
    once (zeros for unrendered Gaussians), optionally straight into the optimizer's gradient slab (`grad_sink`).

Per frame at C3 that is: attribute fwd -> preprocess ... blend -> blend bwd -> per-Gaussian bwd (visible rows) ->
attribute bwd; no torch kernels except the caller's loss.  Only the mesh-bound Gaussians are covered (no background
Gaussians: `render_mask`'s trailing ones, frosting_model.py:1573-1576, need the plain path).
"""
import ctypes as C

import torch

from . import _lib
from . import rasterizer as R
from ._lib import FrostingParams, FrostingGrads


def _p(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


class _FrostingRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bary_logits, opacity_logits, log_scales, quats, sh_dc, sh_rest, inner, outer, cells, faces,
                face_visible, rs, sink):
        if not bary_logits.is_cuda:
            raise RuntimeError("frosting_b200 runs on CUDA tensors only (no CPU fallback)")
        dev = bary_logits.device
        t = [x.contiguous() for x in (bary_logits, opacity_logits, log_scales, quats, sh_dc, sh_rest, inner, outer)]
        bary_logits, opacity_logits, log_scales, quats, sh_dc, sh_rest, inner, outer = t
        cells = cells.to(device=dev, dtype=torch.int64).contiguous()
        faces = faces.to(device=dev, dtype=torch.int32).contiguous()
        fv = None if face_visible is None else face_visible.to(device=dev, dtype=torch.uint8).contiguous()
        P, Rn = bary_logits.shape[0], sh_rest.shape[1]
        fp = FrostingParams(P=P, n_verts=inner.shape[0], n_faces=faces.shape[0], sh_rest=Rn,
                            d_bary_logits=_p(bary_logits), d_cells=_p(cells), d_faces=_p(faces),
                            d_inner_verts=_p(inner), d_outer_verts=_p(outer), d_opacity_logits=_p(opacity_logits),
                            d_log_scales=_p(log_scales), d_quats=_p(quats), d_sh_dc=_p(sh_dc), d_sh_rest=_p(sh_rest),
                            d_mask=None, d_face_visible=_p(fv), d_radii=None)
        o = dict(dtype=torch.float32, device=dev)
        # rows of occluded Gaussians are never read by preprocess (same lookup): plain empty allocations
        means3D, opac = torch.empty((P, 3), **o), torch.empty((P, 1), **o)
        scales, rots, shs = torch.empty((P, 3), **o), torch.empty((P, 4), **o), torch.empty((P, Rn + 1, 3), **o)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fb200_frosting_attributes(
                C.byref(fp), _p(means3D), _p(opac), _p(scales), _p(rots), _p(shs),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        e = torch.Tensor([])
        color, radii, call, _ = R._launch_forward(
            means3D, shs, e, opac, scales, rots, e, rs, None, want_backward=True,
            face_visibility=(fv, cells) if fv is not None else None)
        ctx.fp, ctx.call, ctx.sink, ctx.rs = fp, call, sink, rs
        ctx.keep = (bary_logits, opacity_logits, log_scales, quats, sh_dc, sh_rest, inner, outer, cells, faces, fv)
        ctx.vert_grad = inner.requires_grad or outer.requires_grad
        ctx.save_for_backward(radii)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, g_color, _):
        (radii,) = ctx.saved_tensors
        bary_logits, opacity_logits, log_scales, quats, sh_dc, sh_rest, inner, outer, cells, faces, fv = ctx.keep
        dev = bary_logits.device
        P, Rn = bary_logits.shape[0], sh_rest.shape[1]
        (dL_dmeans2D, _c, dL_dopacity, dL_dmeans3D, _cov, dL_dsh, dL_dscales, dL_drotations, _e) = R._launch_backward(
            ctx.call, radii, g_color, sparse_rows=True)
        o = dict(dtype=torch.float32, device=dev)
        k = ctx.sink
        if k is not None:
            if hasattr(k, "sink_once"):
                k.sink_once(("bary_logits", "opacity_logits", "log_scales", "quats", "sh_dc", "sh_rest"))
            d_bary, d_op, d_ls = k["bary_logits"], k["opacity_logits"], k["log_scales"]
            d_q, d_dc, d_rest = k["quats"], k["sh_dc"], k["sh_rest"]
            for tt, ref in ((d_bary, bary_logits), (d_op, opacity_logits), (d_ls, log_scales), (d_q, quats),
                            (d_dc, sh_dc), (d_rest, sh_rest)):
                if tt.numel() != ref.numel() or not tt.is_contiguous() or tt.dtype != torch.float32 or tt.device != dev:
                    raise RuntimeError("grad_sink tensors must be contiguous fp32 CUDA tensors shaped like the parameters")
        else:
            d_bary, d_op, d_ls = torch.empty((P, 6), **o), torch.empty((P,), **o), torch.empty((P, 3), **o)
            d_q, d_dc, d_rest = torch.empty((P, 4), **o), torch.empty((P, 1, 3), **o), torch.empty((P, Rn, 3), **o)
        d_in = torch.empty_like(inner) if ctx.vert_grad else None
        d_out = torch.empty_like(outer) if ctx.vert_grad else None
        grads = FrostingGrads(d_bary_logits=_p(d_bary), d_inner_verts=_p(d_in), d_outer_verts=_p(d_out),
                              d_opacity_logits=_p(d_op), d_log_scales=_p(d_ls), d_quats=_p(d_q), d_sh_dc=_p(d_dc),
                              d_sh_rest=_p(d_rest))
        fp = ctx.fp
        fp.d_radii = _p(radii)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fb200_frosting_attributes_backward(
                C.byref(fp), _p(dL_dmeans3D), _p(dL_dopacity), _p(dL_dscales), _p(dL_drotations), _p(dL_dsh),
                C.byref(grads), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        if k is not None:
            return (None, None, None, None, None, None, d_in, d_out, None, None, None, None, None)
        return (d_bary, d_op.view_as(opacity_logits), d_ls, d_q, d_dc, d_rest, d_in, d_out, None, None, None, None, None)


def frosting_render(params, mesh, raster_settings, face_visible=None, grad_sink=None):
    """params / mesh: the dicts of `scenes.frosting_layer`; `face_visible` [F]: the prepass's visible-face marks for this
    camera (None: no occlusion culling).  Returns (color [3,H,W], radii [P]); differentiable w.r.t. the parameters and
    the shell vertices."""
    return _FrostingRender.apply(params["bary_logits"], params["opacity_logits"], params["log_scales"], params["quats"],
                                 params["sh_dc"], params["sh_rest"], mesh["inner"], mesh["outer"], mesh["cells"],
                                 mesh["faces"], face_visible, raster_settings, grad_sink)
