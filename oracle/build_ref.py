"""Build the UNMODIFIED reference rasterizer (test infrastructure only).

Compiles the vendored `diff-gaussian-rasterization` sources *where they lie* under
/root/reference (nothing is copied into this repository) into `oracle/_ref/ref_dgr_C.so`
for sm_100a.  `oracle/_ref/` is git-ignored but travels to the GPU box with gpurun.

Recipe: force-include <cstdint> (gcc 13 needs it for
cuda_rasterizer/rasterizer_impl.h:60-61), include the vendored glm, default -O3,
NO fast-math (as DGR/setup.py:20-29).

Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may
import what this script produces; the product (frosting_b200/) never does.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
DGR = "/root/reference/gaussian_splatting/submodules/diff-gaussian-rasterization"
SO = os.path.join(OUT, "ref_dgr_C.so")


def build(verbose: bool = False) -> str:
    if os.path.exists(SO):
        return SO
    if not os.path.isdir(DGR):
        raise FileNotFoundError(
            f"{DGR} not present (GPU box?) and {SO} was not prebuilt")
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "5")
    from torch.utils.cpp_extension import load

    srcs = [
        f"{DGR}/cuda_rasterizer/rasterizer_impl.cu",
        f"{DGR}/cuda_rasterizer/forward.cu",
        f"{DGR}/cuda_rasterizer/backward.cu",
        f"{DGR}/rasterize_points.cu",
        f"{DGR}/ext.cpp",
    ]
    load(
        name="ref_dgr_C",
        sources=srcs,
        extra_include_paths=[f"{DGR}/third_party/glm", DGR],
        extra_cuda_cflags=["-include", "cstdint"],
        extra_cflags=["-include", "cstdint"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=True,
    )
    return SO


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
