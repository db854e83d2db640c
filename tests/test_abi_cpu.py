"""The C-ABI library loads without a GPU and exports exactly what include/frosting_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from frosting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "frosting_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED) == names, "frosting_b200/_lib.py:EXPORTED out of sync with the header"
    assert L.fb200_abi_version() == _lib.ABI_VERSION == 4


def test_workspace_size_queries_are_host_only_and_monotonic():
    L = _lib.lib()
    assert L.fb200_geom_bytes(0) >= 128
    g1, g2 = L.fb200_geom_bytes(1000), L.fb200_geom_bytes(2000)
    assert g2 > g1 >= 1000 * 109
    i1 = L.fb200_image_bytes(1920, 1080)
    assert i1 >= 1920 * 1080 * 12 + 8160 * 8
    assert L.fb200_image_bytes(1921, 1080) > i1
    b = L.fb200_binning_bytes(10_000)
    assert 10_000 * 52 <= b <= 10_000 * 52 + 1024      # 20 B of keys + 8 sub-tile hit lists of 4 B per instance


def test_layout_offsets_are_aligned_and_ordered():
    L = _lib.lib()
    lay = _lib.Layout()
    assert L.fb200_get_layout(1000, 256, 128, 5000, C.byref(lay)) == 0
    offs = [getattr(lay, n) for n, _ in _lib.Layout._fields_]
    assert all(o % 128 == 0 for o in offs)
    assert lay.geom_rec < lay.geom_depth < lay.geom_rect < lay.geom_clamped
    assert lay.img_final_T < lay.img_n_contrib < lay.img_ranges < lay.img_tile_count
    assert lay.bin_point_list < lay.bin_keys


def test_argument_errors_are_reported_without_touching_cuda():
    L = _lib.lib()
    lay = _lib.Layout()
    assert L.fb200_get_layout(-1, 256, 128, 0, C.byref(lay)) == -1
    assert b"layout" in L.fb200_last_error()
    prm = _lib.Params(P=10, sh_degree=0, sh_coeffs=0, image_width=64, image_height=64, tanfovx=0.5, tanfovy=0.5,
                      scale_modifier=1.0, prefiltered=0, debug=0)
    inp = _lib.Inputs()      # all NULL
    ws = _lib.Workspace()
    rc = L.fb200_forward(C.byref(prm), C.byref(inp), C.byref(ws), None, None, None)
    assert rc == -1 and b"missing" in L.fb200_last_error()
    # exactly-one-of checks carry the reference's messages (DGR/diff_gaussian_rasterization/__init__.py:191-195)
    dummy = C.c_void_p(0x1000)
    inp = _lib.Inputs(d_background=dummy, d_means3D=dummy, d_opacities=dummy, d_viewmatrix=dummy,
                      d_projmatrix=dummy, d_campos=dummy)
    rc = L.fb200_forward(C.byref(prm), C.byref(inp), C.byref(ws), None, None, None)
    assert rc == -1 and b"excatly one of either SHs or precomputed colors" in L.fb200_last_error()
    inp.d_colors_precomp = dummy
    rc = L.fb200_forward(C.byref(prm), C.byref(inp), C.byref(ws), None, None, None)
    assert rc == -1 and b"exactly one of either scale/rotation pair" in L.fb200_last_error()
    with pytest.raises(_lib.Fb200Error):
        _lib.check(rc)
    # frosting mode: the parameter block replaces the attribute pointers, it does not coexist with them
    d = dummy
    fp = _lib.FrostingParams(P=10, n_verts=4, n_faces=2, sh_rest=15, d_bary_logits=d, d_cells=d, d_faces=d, d_inner_verts=d,
                             d_outer_verts=d, d_opacity_logits=d, d_log_scales=d, d_quats=d, d_sh_dc=d, d_sh_rest=d)
    prm.sh_degree, prm.sh_coeffs = 3, 16
    inp = _lib.Inputs(d_background=d, d_means3D=d, d_viewmatrix=d, d_projmatrix=d, d_campos=d, frosting=C.addressof(fp))
    rc = L.fb200_forward(C.byref(prm), C.byref(inp), C.byref(ws), None, None, None)
    assert rc == -1 and b"frosting mode: the attribute / mask pointers" in L.fb200_last_error()
    inp.d_means3D = None
    prm.sh_coeffs = 9
    rc = L.fb200_forward(C.byref(prm), C.byref(inp), C.byref(ws), None, None, None)
    assert rc == -1 and b"do not match the parameter block" in L.fb200_last_error()
    prm.sh_coeffs = 16
    rc = L.fb200_forward(C.byref(prm), C.byref(inp), C.byref(ws), None, None, None)
    assert rc == -1 and b"workspace pointers missing" in L.fb200_last_error()      # got past the frosting checks


def test_profile_and_launch_counter_entry_points():
    assert _lib.kernel_launches() >= 0
    _lib.profile_enable(True)
    t = _lib.profile_read()
    assert set(t) == set(_lib.STAGES) and all(v == -1.0 for v in t.values())   # nothing recorded without a GPU
    _lib.profile_enable(False)
