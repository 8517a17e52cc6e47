"""Build recipe for the UNMODIFIED reference CUDA extensions (test infrastructure only).

Compiles the reference's own sources *where they lie* under /root/reference into
oracle/_ref/ (git-ignored, travels to the GPU box with the gpurun snapshot).  Nothing from
/root/reference is copied into the repo.  The two extensions are used ONLY by tests/, by
`__graft_entry__.smoke()` and by `bench.py --impl reference` as the parity pin / baseline:

  ref_dgr_C : Edit_core/thirdparties/diff-gaussian-rasterization (ext.cpp:15-19 entry points)
  ref_knn_C : Edit_core/thirdparties/simple-knn                  (ext.cpp:15-17 distCUDA2)

gcc 13 needs `-include cstdint` (rasterizer_impl.h:24 uses std::uintptr_t without the header)
and simple_knn.cu needs `-include cfloat` (FLT_MAX, simple_knn.cu:90,154).  No source edits.
"""
import os
import sys

REF = os.environ.get("TGR_REFERENCE_ROOT", "/root/reference")
DGR = os.path.join(REF, "Edit_core/thirdparties/diff-gaussian-rasterization")
KNN = os.path.join(REF, "Edit_core/thirdparties/simple-knn")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def available() -> bool:
    return os.path.isdir(DGR) and os.path.isdir(KNN)


def build(verbose: bool = False) -> None:
    if not available():
        raise RuntimeError("reference sources not found under %s" % REF)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    d = os.path.join(OUT, "dgr")
    os.makedirs(d, exist_ok=True)
    load(
        name="ref_dgr_C",
        sources=[
            os.path.join(DGR, "cuda_rasterizer/rasterizer_impl.cu"),
            os.path.join(DGR, "cuda_rasterizer/forward.cu"),
            os.path.join(DGR, "cuda_rasterizer/backward.cu"),
            os.path.join(DGR, "rasterize_points.cu"),
            os.path.join(DGR, "ext.cpp"),
        ],
        extra_include_paths=[DGR, os.path.join(DGR, "third_party/glm")],
        extra_cuda_cflags=ARCH + ["-include", "cstdint", "-O3"],
        extra_cflags=["-O3"],
        build_directory=d,
        is_python_module=False,
        verbose=verbose,
    )
    k = os.path.join(OUT, "knn")
    os.makedirs(k, exist_ok=True)
    load(
        name="ref_knn_C",
        sources=[
            os.path.join(KNN, "spatial.cu"),
            os.path.join(KNN, "simple_knn.cu"),
            os.path.join(KNN, "ext.cpp"),
        ],
        extra_include_paths=[KNN],
        extra_cuda_cflags=ARCH + ["-include", "cfloat", "-O3"],
        extra_cflags=["-O3"],
        build_directory=k,
        is_python_module=False,
        verbose=verbose,
    )


if __name__ == "__main__":
    build(verbose="-v" in sys.argv)
    print("built:", [os.path.join(OUT, p) for p in ("dgr/ref_dgr_C.so", "knn/ref_knn_C.so")])
