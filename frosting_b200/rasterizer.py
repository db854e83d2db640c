"""Python surface of the B200 rasterizer -- a drop-in for `diff_gaussian_rasterization`.

Mirrors DGR/diff_gaussian_rasterization/__init__.py (DGR = gaussian_splatting/submodules/
diff-gaussian-rasterization in the reference):
  GaussianRasterizationSettings  <- __init__.py:157-169 (same 12 fields, same order)
  GaussianRasterizer             <- __init__.py:171-220 (same forward signature, same exceptions,
                                    same (color[3,H,W], radii[P]) return, markVisible)
  _RasterizeGaussians            <- __init__.py:44-155  (same 8 gradients in the same order)
The native side is reached through the C ABI in include/frosting_b200.h via ctypes; torch only owns
the tensors and the stream.

One extension: `GaussianRasterizer.forward(..., visibility_mask=None)` takes the per-Gaussian
occlusion mask (frosting_model.py:1564-1576) and drops masked Gaussians inside the preprocess kernel
instead of boolean-gathering every attribute tensor (frosting_model.py:1578-1586).
"""
from typing import NamedTuple, Optional
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import Params, Inputs, Workspace, Grads, Layout


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


_last = {"num_rendered": 0}
NO_CULL = False   # test hook: disable sub-tile culling (debug bit 1) to prove it is output-neutral


_pinned = {}
_size_cache = {}


def _sizes(L, P, W, H):
    """Workspace sizes rounded up to 512 bytes (so the carved sub-buffers keep torch's base alignment)."""
    k = (P, W, H)
    v = _size_cache.get(k)
    if v is None:
        r = lambda n: (int(n) + 511) // 512 * 512
        v = (r(L.fb200_geom_bytes(P)), r(L.fb200_image_bytes(W, H)))
        if len(_size_cache) > 64:
            _size_cache.clear()
        _size_cache[k] = v
    return v


_capacity_hint = {}     # (device, P, W, H) -> binning capacity to pre-allocate


def _pinned_status(device):
    # one pinned 32-byte mailbox per device: cudaHostAlloc per call would cost more than the frame
    t = _pinned.get(device.index)
    if t is None:
        t = torch.empty((_lib.FB200_STATUS_WORDS,), dtype=torch.int32, pin_memory=True)
        _pinned[device.index] = t
    return t


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, device, name: str) -> torch.Tensor:
    if t.numel() == 0:
        return t
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.device != device:
        t = t.to(device)
    return t if t.is_contiguous() else t.contiguous()


class _Call:
    """Everything one forward needs again in backward (kept alive by the autograd ctx)."""
    __slots__ = ("prm", "inp", "ws", "tensors", "geom", "image", "binning", "status", "capacity",
                 "num_rendered", "device", "extra")


def _launch_forward(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    rs: GaussianRasterizationSettings, visibility, extra_features=None, extra_bg=None):
    L = _lib.lib()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")   # rasterize_points.cu:57-59
    if not means3D.is_cuda:
        raise RuntimeError("frosting_b200 runs on CUDA tensors only (no CPU fallback)")
    device = means3D.device
    P = means3D.size(0)
    H, W = int(rs.image_height), int(rs.image_width)

    means3D = _f32c(means3D, device, "means3D")
    sh = _f32c(sh, device, "shs")
    colors_precomp = _f32c(colors_precomp, device, "colors_precomp")
    opacities = _f32c(opacities, device, "opacities")
    scales = _f32c(scales, device, "scales")
    rotations = _f32c(rotations, device, "rotations")
    cov3Ds_precomp = _f32c(cov3Ds_precomp, device, "cov3D_precomp")
    bg = _f32c(rs.bg, device, "bg")
    view = _f32c(rs.viewmatrix, device, "viewmatrix")
    proj = _f32c(rs.projmatrix, device, "projmatrix")
    campos = _f32c(rs.campos, device, "campos")
    vis = None
    if visibility is not None:
        vis = visibility.to(device=device, dtype=torch.uint8).contiguous() if visibility.dtype != torch.uint8 \
            else visibility.to(device).contiguous()
        if vis.numel() != P:
            raise RuntimeError("visibility_mask must have one entry per Gaussian")

    M = sh.size(1) if sh.numel() != 0 else 0   # rasterize_points.cu:84-87

    # row f4: extra feature channels blended in the same traversal
    extra = None
    if extra_features is not None:
        extra_features = _f32c(extra_features, device, "extra_features")
        if extra_features.dim() != 2 or extra_features.size(0) != P or not (1 <= extra_features.size(1) <= 3):
            raise RuntimeError("extra_features must have dimensions (num_points, 1..3)")
        E = extra_features.size(1)
        extra_bg = torch.zeros(E, dtype=torch.float32, device=device) if extra_bg is None \
            else _f32c(extra_bg.reshape(-1), device, "extra_background")
        if extra_bg.numel() != E:
            raise RuntimeError("extra_background must have one value per extra channel")

    prm = Params(P=P, sh_degree=int(rs.sh_degree), sh_coeffs=int(M), image_width=W, image_height=H,
                 tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy),
                 scale_modifier=float(rs.scale_modifier), prefiltered=int(bool(rs.prefiltered)),
                 debug=int(bool(rs.debug)) | (2 if NO_CULL else 0), extra=None)
    inp = Inputs(d_background=_ptr(bg), d_means3D=_ptr(means3D), d_shs=_ptr(sh),
                 d_colors_precomp=_ptr(colors_precomp), d_opacities=_ptr(opacities), d_scales=_ptr(scales),
                 d_rotations=_ptr(rotations), d_cov3D_precomp=_ptr(cov3Ds_precomp),
                 d_viewmatrix=_ptr(view), d_projmatrix=_ptr(proj), d_campos=_ptr(campos), d_visibility=_ptr(vis))

    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        out_color = torch.empty((3, H, W), dtype=torch.float32, device=device)
        radii = torch.empty((P,), dtype=torch.int32, device=device)
        out_extra = None
        if extra_features is not None:
            out_extra = torch.empty((extra_features.size(1), H, W), dtype=torch.float32, device=device)
            if P == 0:
                out_extra.copy_(extra_bg.view(-1, 1, 1).expand_as(out_extra))
            extra = _lib.Extra(channels=extra_features.size(1), d_features=_ptr(extra_features),
                               d_background=_ptr(extra_bg), d_out=_ptr(out_extra), d_dL_dout=None, d_dL_dfeatures=None)
            prm.extra = C.addressof(extra)
        gb, ib = _sizes(L, P, W, H)
        slab = torch.empty((gb + ib + 128,), dtype=torch.uint8, device=device)   # one allocator call
        geom, image = slab[:gb], slab[gb:gb + ib]
        status = slab[gb + ib:gb + ib + 4 * _lib.FB200_STATUS_WORDS].view(torch.int32)
        status_host = _pinned_status(device)

        # Phase 1: preprocess + tile scan.  The exact instance count R comes back through a pinned
        # mailbox; waiting for it covers ~0.1 ms of GPU work (the reference blocks at the same point,
        # rasterizer_impl.cu:280-281).  Phase 2 is then launched with an exactly sized binning buffer and
        # the caller's loss/backward launches queue up behind it while the GPU is busy.
        ws = Workspace(d_geom=geom.data_ptr(), geom_bytes=geom.numel(),
                       d_image=image.data_ptr(), image_bytes=image.numel(),
                       d_binning=None, binning_bytes=0, binning_capacity=0, d_status=status.data_ptr(),
                       # a backward will follow: clear its accumulators while the host waits for R (below)
                       acc_zeroed_by_forward=1 if torch.is_grad_enabled() else 0)
        rptr = C.c_void_p(radii.data_ptr()) if P > 0 else None
        sptr = C.c_void_p(stream.cuda_stream)
        _lib.check(L.fb200_forward_geometry(C.byref(prm), C.byref(inp), C.byref(ws), rptr, sptr))
        status_host.copy_(status, non_blocking=True)
        # the binning buffer is allocated BEFORE the wait from the last count seen for this problem size, so that the
        # host's critical path after the wait is one comparison and one library call
        hint_key = (device.index, P, W, H)
        capacity = _capacity_hint.get(hint_key, 0)
        binning = torch.empty((L.fb200_binning_bytes(capacity),), dtype=torch.uint8, device=device) if capacity else None
        stream.synchronize()
        num_rendered = int(status_host[_lib.ST_NUM_RENDERED])
        if num_rendered > capacity or binning is None:
            capacity = max(num_rendered, 1)
            binning = torch.empty((L.fb200_binning_bytes(capacity),), dtype=torch.uint8, device=device)
        _capacity_hint[hint_key] = max(int(num_rendered * 1.05) + 1024, 1)
        ws.d_binning, ws.binning_bytes, ws.binning_capacity = binning.data_ptr(), binning.numel(), capacity
        ws.h_status = status_host.data_ptr()      # the mailbox holds this frame's status words until the next forward
        _lib.check(L.fb200_forward_raster(C.byref(prm), C.byref(inp), C.byref(ws),
                                          C.c_void_p(out_color.data_ptr()), rptr, sptr))
        _last["num_rendered"] = num_rendered

    call = _Call()
    call.prm, call.inp, call.ws = prm, inp, ws
    call.tensors = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, bg, view, proj,
                    campos, vis, extra_features, extra_bg)
    call.extra = extra
    call.geom, call.image, call.binning, call.status = geom, image, binning, status
    call.capacity, call.num_rendered, call.device = capacity, num_rendered, device
    return out_color, radii, call, out_extra


def _launch_backward(call: "_Call", radii, grad_out_color, grad_out_extra=None):
    L = _lib.lib()
    prm = call.prm
    P, M = prm.P, prm.sh_coeffs
    device = call.device
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        g = grad_out_color
        if g.dtype != torch.float32:
            g = g.float()
        g = g.contiguous()
        # one allocation carved into the gradient tensors the caller can use (each offset a multiple of 4 floats);
        # dL/dcolors on the SH path, dL/dcov3D on the scale/rotation path and dL/dscales, dL/drotations on the
        # precomputed-covariance path are never returned (`backward` below) and are not materialised
        has_sh, has_col = call.inp.d_shs is not None, call.inp.d_colors_precomp is not None
        has_cov = call.inp.d_cov3D_precomp is not None
        n = [3 * P, 3 * P, 3 * P if has_col else 0, P, 6 * P if has_cov else 0, 3 * M * P if has_sh else 0,
             0 if has_cov else 3 * P, 0 if has_cov else 4 * P]
        offs, tot = [], 0
        for c in n:
            offs.append(tot)
            tot += (c + 3) // 4 * 4
        gslab = torch.empty((max(tot, 1),), dtype=torch.float32, device=device)
        cut = lambda k, shape: gslab[offs[k]:offs[k] + n[k]].view(shape) if n[k] or P == 0 else None
        dL_dmeans3D, dL_dmeans2D, dL_dcolors = cut(0, (P, 3)), cut(1, (P, 3)), cut(2, (P, 3))
        dL_dopacity, dL_dcov3D, dL_dsh = cut(3, (P, 1)), cut(4, (P, 6)), cut(5, (P, M, 3))
        dL_dscales, dL_drotations = cut(6, (P, 3)), cut(7, (P, 4))
        dL_dextra = None
        if call.extra is not None:
            E = call.extra.channels
            ge = grad_out_extra
            ge = torch.zeros((E, prm.image_height, prm.image_width), dtype=torch.float32, device=device) if ge is None \
                else ge.float().contiguous()
            dL_dextra = torch.empty((P, E), dtype=torch.float32, device=device)
            call.extra.d_dL_dout, call.extra.d_dL_dfeatures = ge.data_ptr(), dL_dextra.data_ptr()
            prm.extra = C.addressof(call.extra)
        grads = Grads(d_dL_dmeans2D=_ptr(dL_dmeans2D), d_dL_dcolors=_ptr(dL_dcolors),
                      d_dL_dopacity=_ptr(dL_dopacity), d_dL_dmeans3D=_ptr(dL_dmeans3D),
                      d_dL_dcov3D=_ptr(dL_dcov3D), d_dL_dsh=_ptr(dL_dsh), d_dL_dscales=_ptr(dL_dscales),
                      d_dL_drotations=_ptr(dL_drotations))
        _lib.check(L.fb200_backward(C.byref(prm), C.byref(call.inp), C.byref(call.ws),
                                    C.c_void_p(radii.data_ptr()) if P > 0 else None,
                                    C.c_void_p(g.data_ptr()), C.byref(grads),
                                    C.c_void_p(stream.cuda_stream)))
        call.ws.acc_zeroed_by_forward = 0      # a second backward over this forward must clear them itself
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dextra


def last_num_rendered() -> int:
    """R (tile instances) of the most recent forward in this process."""
    return _last["num_rendered"]


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, visibility_mask=None, extra_features=None, extra_background=None):
    out = _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                    cov3Ds_precomp, raster_settings, visibility_mask, extra_features, extra_background)
    return out if extra_features is not None else out[:2]


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, visibility_mask=None, extra_features=None, extra_background=None):
        args = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                visibility_mask, extra_features, extra_background)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args[:7])   # copy before they can be corrupted
            try:
                color, radii, call, extra = _launch_forward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            color, radii, call, extra = _launch_forward(*args)
        ctx.call = call
        ctx.num_rendered = call.num_rendered
        ctx.raster_settings = raster_settings
        ctx.present = tuple(t.numel() != 0 for t in (sh, colors_precomp, scales, rotations, cov3Ds_precomp))
        ctx.save_for_backward(radii)
        ctx.mark_non_differentiable(radii)
        if extra is None:
            extra = color.new_empty(0)        # placeholder third output (autograd wants tensors)
        return color, radii, extra

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_extra=None):
        (radii,) = ctx.saved_tensors
        call = ctx.call
        if ctx.raster_settings.debug:
            try:
                out = _launch_backward(call, radii, grad_out_color, grad_out_extra)
            except Exception as ex:
                torch.save(cpu_deep_copy_tuple(call.tensors[:7]) + (grad_out_color.cpu().clone(),), "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _launch_backward(call, radii, grad_out_color, grad_out_extra)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
         grad_scales, grad_rotations, grad_extra) = out
        has_sh, has_col, has_s, has_r, has_cov = ctx.present
        # same order as DGR/diff_gaussian_rasterization/__init__.py:143-153
        return (
            grad_means3D,
            grad_means2D,
            grad_sh if has_sh else None,
            grad_colors_precomp if has_col else None,
            grad_opacities,
            grad_scales if has_s else None,
            grad_rotations if has_r else None,
            grad_cov3Ds_precomp if has_cov else None,
            None,
            None,
            grad_extra,
            None,
        )


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points passing the near-plane test (checkFrustum, rasterizer_impl.cu:54-66)."""
        with torch.no_grad():
            rs = self.raster_settings
            if not positions.is_cuda:
                raise RuntimeError("frosting_b200 runs on CUDA tensors only (no CPU fallback)")
            device = positions.device
            pos = _f32c(positions, device, "positions")
            view = _f32c(rs.viewmatrix, device, "viewmatrix")
            proj = _f32c(rs.projmatrix, device, "projmatrix")
            P = pos.size(0)
            present = torch.empty((P,), dtype=torch.bool, device=device)
            with torch.cuda.device(device):
                _lib.check(_lib.lib().fb200_mark_visible(
                    P, _ptr(pos), _ptr(view), _ptr(proj), _ptr(present),
                    C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, visibility_mask=None, extra_features=None, extra_background=None):
        """Reference surface (`__init__.py:187-220`) plus two opt-in extensions: `visibility_mask` (row a19) and
        `extra_features` [P, 1..3] (row f4) -- blended with the colour's weights in the same traversal; when given, a
        third output `[E, H, W]` is returned (what a second call with `colors_precomp=extra_features` and
        `bg=extra_background` would return, sugar_model.py:2343-2387)."""
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings, visibility_mask, extra_features, extra_background)


# ---- introspection for parity tests ----------------------------------------------------------------------
def forward_with_state(raster_settings, means3D, opacities, shs=None, colors_precomp=None, scales=None,
                       rotations=None, cov3D_precomp=None, visibility_mask=None):
    """Run the forward and expose the internal arrays (no autograd): used by tests to compare depth
    bits, rects, records, ranges and the sorted point list with the reference's buffers."""
    e = torch.Tensor([])
    with torch.no_grad():
        color, radii, call, _ = _launch_forward(
            means3D, e if shs is None else shs, e if colors_precomp is None else colors_precomp, opacities,
            e if scales is None else scales, e if rotations is None else rotations,
            e if cov3D_precomp is None else cov3D_precomp, raster_settings, visibility_mask)
    P, W, H = call.prm.P, call.prm.image_width, call.prm.image_height
    lay = Layout()
    _lib.check(_lib.lib().fb200_get_layout(P, W, H, call.capacity, C.byref(lay)))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    R = call.num_rendered

    def view(buf, off, nbytes, dtype):
        base = (-buf.data_ptr()) % 128   # the library aligns the base to 128 bytes
        return buf[base + off: base + off + nbytes].view(dtype)

    st = dict(
        color=color, radii=radii, num_rendered=R, call=call,
        rec=view(call.geom, lay.geom_rec, P * 48, torch.float32).view(P, 12),
        depth=view(call.geom, lay.geom_depth, P * 4, torch.float32),
        rect=view(call.geom, lay.geom_rect, P * 8, torch.int32).view(P, 2),
        clamped=view(call.geom, lay.geom_clamped, P, torch.uint8),
        final_T=view(call.image, lay.img_final_T, W * H * 4, torch.float32).view(H, W),
        n_contrib=view(call.image, lay.img_n_contrib, W * H * 4, torch.int32).view(H, W),
        ranges=view(call.image, lay.img_ranges, T * 8, torch.int32).view(T, 2),
        tile_count=view(call.image, lay.img_tile_count, T * 4, torch.int32),
        point_list=view(call.binning, lay.bin_point_list, R * 4, torch.int32),
        keys=view(call.binning, lay.bin_keys, R * 8, torch.int64),
    )
    return st
