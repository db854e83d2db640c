"""ctypes loader for the C oracle (oracle/raster_oracle.c) with numpy-in / numpy-out wrappers.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Builds the shared object on first use with gcc
(-ffp-contract=off so that only the explicit fmaf() calls fuse).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "raster_oracle.c")
_SO = os.path.join(_HERE, "_build", "libraster_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        r = subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "_build/libraster_oracle.so"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_bin.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Camera:
    """The per-call camera record (fields of GaussianRasterizationSettings as numpy)."""

    def __init__(self, rs):
        self.W, self.H = int(rs.image_width), int(rs.image_height)
        self.tanx, self.tany = float(np.float32(rs.tanfovx)), float(np.float32(rs.tanfovy))
        self.view = _f32(rs.viewmatrix.detach().cpu().numpy()).reshape(16)
        self.proj = _f32(rs.projmatrix.detach().cpu().numpy()).reshape(16)
        self.campos = _f32(rs.campos.detach().cpu().numpy()).reshape(3)
        self.bg = _f32(rs.bg.detach().cpu().numpy()).reshape(3)
        self.D = int(rs.sh_degree)
        self.mod = float(np.float32(rs.scale_modifier))


def preprocess(cam, means, opac, shs=None, colors_precomp=None, scales=None, rots=None, cov3D_precomp=None,
               visibility=None):
    L = lib()
    means, opac = _f32(means), _f32(opac).reshape(-1)
    shs, colors_precomp, scales, rots, cov3D_precomp = map(_f32, (shs, colors_precomp, scales, rots, cov3D_precomp))
    P = means.shape[0]
    M = 0 if shs is None else shs.shape[1]
    vis = None if visibility is None else np.ascontiguousarray(visibility, dtype=np.uint8)
    out = dict(radii=np.zeros(P, np.int32), xy=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
               cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
               rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8),
               tiles_touched=np.zeros(P, np.uint32), rect=np.zeros((P, 4), np.uint32))
    L.oracle_preprocess(C.c_int(P), C.c_int(cam.D), C.c_int(M), _p(means), _p(scales), C.c_float(cam.mod), _p(rots),
                        _p(opac), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(cam.view), _p(cam.proj),
                        _p(cam.campos), C.c_int(cam.W), C.c_int(cam.H), C.c_float(cam.tanx), C.c_float(cam.tany),
                        _p(vis), _p(out["radii"]), _p(out["xy"]), _p(out["depths"]), _p(out["cov3D"]),
                        _p(out["conic_opacity"]), _p(out["rgb"]), _p(out["clamped"]), _p(out["tiles_touched"]),
                        _p(out["rect"]))
    if colors_precomp is not None:
        out["rgb"] = colors_precomp.copy()
    if cov3D_precomp is not None:
        out["cov3D"] = cov3D_precomp.copy()
    return out


def bin_and_sort(cam, pre):
    L = lib()
    P = pre["radii"].shape[0]
    T = ((cam.W + 15) // 16) * ((cam.H + 15) // 16)
    R = int(L.oracle_bin(C.c_int(P), C.c_int(cam.W), C.c_int(cam.H), _p(pre["radii"]), _p(pre["rect"]),
                         _p(pre["depths"]), None, None, None))
    keys = np.zeros(max(R, 1), np.uint64)
    plist = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((T, 2), np.uint32)
    L.oracle_bin(C.c_int(P), C.c_int(cam.W), C.c_int(cam.H), _p(pre["radii"]), _p(pre["rect"]), _p(pre["depths"]),
                 _p(keys), _p(plist), _p(ranges))
    return dict(num_rendered=R, keys=keys[:R], point_list=plist[:R], ranges=ranges)


def render_fwd(cam, pre, binned):
    L = lib()
    N = cam.W * cam.H
    out = dict(final_T=np.zeros(N, np.float32), n_contrib=np.zeros(N, np.uint32),
               color=np.zeros((3, cam.H, cam.W), np.float32))
    L.oracle_render_fwd(C.c_int(cam.W), C.c_int(cam.H), _p(binned["ranges"]), _p(binned["point_list"]), _p(pre["xy"]),
                        _p(pre["rgb"]), _p(pre["conic_opacity"]), _p(cam.bg), _p(out["final_T"]), _p(out["n_contrib"]),
                        _p(out["color"]))
    out["final_T"] = out["final_T"].reshape(cam.H, cam.W)
    out["n_contrib"] = out["n_contrib"].reshape(cam.H, cam.W)
    return out


def forward(rs, means, opac, **kw):
    cam = Camera(rs)
    pre = preprocess(cam, means, opac, **kw)
    binned = bin_and_sort(cam, pre)
    img = render_fwd(cam, pre, binned)
    return dict(cam=cam, pre=pre, binned=binned, **img)


def backward(rs, fwd, means, dL_dpix, shs=None, colors_precomp=None, scales=None, rots=None, cov3D_precomp=None):
    """Full backward; returns the reference's 8 gradient tensors as numpy (fp32), accumulated in fp64."""
    L = lib()
    cam, pre, binned = fwd["cam"], fwd["pre"], fwd["binned"]
    means = _f32(means)
    P = means.shape[0]
    dL_dpix = _f32(dL_dpix)
    g2 = np.zeros((P, 2), np.float64); gc = np.zeros((P, 3), np.float64)
    go = np.zeros(P, np.float64); gcol = np.zeros((P, 3), np.float64)
    L.oracle_render_bwd(C.c_int(P), C.c_int(cam.W), C.c_int(cam.H), _p(binned["ranges"]), _p(binned["point_list"]),
                        _p(cam.bg), _p(pre["xy"]), _p(pre["conic_opacity"]), _p(pre["rgb"]),
                        _p(np.ascontiguousarray(fwd["final_T"].reshape(-1))),
                        _p(np.ascontiguousarray(fwd["n_contrib"].reshape(-1))), _p(dL_dpix),
                        _p(g2), _p(gc), _p(go), _p(gcol))
    shs, scales, rots = _f32(shs), _f32(scales), _f32(rots)
    M = 0 if shs is None else shs.shape[1]
    g2f, gcf, gcolf = g2.astype(np.float32), gc.astype(np.float32), gcol.astype(np.float32)
    out = dict(means3D=np.zeros((P, 3), np.float32), cov3D=np.zeros((P, 6), np.float32),
               sh=np.zeros((P, M, 3), np.float32), scales=np.zeros((P, 3), np.float32),
               rotations=np.zeros((P, 4), np.float32))
    L.oracle_geom_bwd(C.c_int(P), C.c_int(cam.D), C.c_int(M), _p(means), _p(pre["radii"]), _p(shs), _p(pre["clamped"]),
                      _p(scales), _p(rots), C.c_float(cam.mod), _p(np.ascontiguousarray(pre["cov3D"])), _p(cam.view),
                      _p(cam.proj), _p(cam.campos), C.c_int(cam.W), C.c_int(cam.H), C.c_float(cam.tanx),
                      C.c_float(cam.tany), _p(g2f), _p(gcf), _p(gcolf), _p(out["means3D"]), _p(out["cov3D"]),
                      _p(out["sh"]) if M else None, _p(out["scales"]), _p(out["rotations"]))
    out["means2D"] = np.concatenate([g2f, np.zeros((P, 1), np.float32)], 1)
    out["colors"] = gcolf
    out["opacities"] = go.astype(np.float32).reshape(P, 1)
    return out


def mark_visible(rs, means):
    cam = Camera(rs)
    means = _f32(means)
    out = np.zeros(means.shape[0], np.uint8)
    lib().oracle_mark_visible(C.c_int(means.shape[0]), _p(means), _p(cam.view), _p(out))
    return out.astype(bool)


def mesh_raster(verts, faces, full_proj, H, W, mark_last_on_bg=False):
    verts = _f32(verts); faces = np.ascontiguousarray(faces, dtype=np.int32)
    m = _f32(full_proj).reshape(16)
    p2f = np.zeros((H, W), np.int32)
    fv = np.zeros(max(faces.shape[0], 1), np.uint8)
    lib().oracle_mesh_raster(C.c_int(verts.shape[0]), C.c_int(faces.shape[0]), _p(verts), _p(faces), _p(m), C.c_int(W),
                             C.c_int(H), _p(p2f), _p(fv), C.c_int(int(mark_last_on_bg)))
    return p2f, fv[:faces.shape[0]].astype(bool)
