"""Golden vectors for Frosting's parameter -> attribute chain, produced by EXECUTING the
reference's own property source.

`frosting_scene/frosting_model.py` cannot be imported here (it needs pytorch3d, open3d, simple_knn), but the properties
that turn the learnable parameters into the rasterizer's inputs are plain torch: their source is cut out of the
reference file with `ast` -- outer_verts / inner_verts (:643-710), shell_cells_verts (:705-710), bary_coords (:713-719),
points (:721-726), strengths (:728-730), sh_coordinates (:732-734), scaling (:736-766), quaternions (:768-799) --
compiled UNMODIFIED into a stub class and run on a small frosting layer with the flags the refinement stage uses
(editable=False, use_softmax_for_bary_coords=True, positions_are_absolute=False, scale_activation=torch.exp, :32).

    python tests/golden/make_frosting_attr_golden.py        # writes tests/golden/frosting_attrs.npz

Stores the parameters, the mesh, the five outputs, cotangents and the autograd gradients w.r.t. every parameter and
the shell vertices.  tests/test_oracle_golden_cpu.py pins `scenes.frosting_attributes` (the restatement the GPU tests use as
their oracle) to these vectors; tests/test_parity_gpu.py pins the fused CUDA kernels to them.
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/frosting_scene/frosting_model.py"
WANT = ("outer_verts", "inner_verts", "shell_cells_verts", "bary_coords", "points", "strengths", "sh_coordinates",
        "scaling", "quaternions")


def reference_properties():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Frosting")
    lines = src.splitlines()
    chunks = []
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in WANT:
            first = min([d.lineno for d in n.decorator_list] + [n.lineno])
            chunks.append("\n".join(lines[first - 1:n.end_lineno]))
    assert len(chunks) == len(WANT), [c.split("def ")[1].split("(")[0] for c in chunks]
    code = "class RefProps:\n" + "\n\n".join(chunks) + "\n"
    ns = {"torch": torch}
    exec(compile(code, REF, "exec"), ns)       # the reference's own statements, line for line
    return ns["RefProps"]


def main():
    from frosting_b200 import scenes
    RefProps = reference_properties()
    P = 1200
    cam = scenes.make_camera(160, 96)
    params, mesh = scenes.frosting_layer(P, cam, 17, n_faces_target=300)
    thickness = 0.02
    normals = (mesh["outer"] - mesh["inner"]) / (2 * thickness)
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    base = mesh["verts"].clone().requires_grad_(True)
    m = RefProps()
    m.editable, m.use_softmax_for_bary_coords, m.positions_are_absolute = False, True, False
    m.scale_activation = torch.exp
    m._shell_base_verts, m.shell_base_normals = base, normals
    m._outer_dist = torch.full((base.shape[0],), thickness)
    m._inner_dist = torch.full((base.shape[0],), -thickness)
    m._shell_base_faces = mesh["faces"].long()
    m._point_cell_indices = mesh["cells"]
    m._bary_coords, m._opacities = leaf["bary_logits"], leaf["opacity_logits"]
    m._sh_coordinates_dc, m._sh_coordinates_rest = leaf["sh_dc"], leaf["sh_rest"]
    m._scales, m._quaternions = leaf["log_scales"], leaf["quats"]
    out = dict(means3D=m.points, opacities=m.strengths, scales=m.scaling, rotations=m.quaternions, shs=m.sh_coordinates)
    g = torch.Generator().manual_seed(5)
    cots = {k: torch.randn(v.shape, generator=g) for k, v in out.items()}
    sum((out[k] * cots[k]).sum() for k in out).backward()
    save = {f"param_{k}": v.detach().numpy() for k, v in leaf.items()}
    save.update({f"mesh_{k}": v.numpy() for k, v in mesh.items()})
    save.update({f"out_{k}": v.detach().numpy() for k, v in out.items()})
    save.update({f"cot_{k}": v.numpy() for k, v in cots.items()})
    save.update({f"grad_{k}": v.grad.numpy() for k, v in leaf.items()})
    # inner = base - t n, outer = base + t n: the gradient w.r.t. the base vertices is the sum of the two shells' gradients
    save["grad_base_verts"] = base.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "frosting_attrs.npz"), **save)
    print({k: v.shape for k, v in save.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
