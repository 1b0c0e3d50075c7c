#!/usr/bin/env python
"""Build the UNMODIFIED reference rasterizer extensions into oracle/_ref/ (test infrastructure only).

This is the recipe SURVEY.md §8(c) verified: the reference's own CUDA sources are compiled *where they lie*
under /root/reference (read-only, never copied into this repo) with torch.utils.cpp_extension.load, the only
compatibility fix being ``-include cstdint`` (gcc 13 needs it for ``uint32_t`` in cuda_rasterizer/rasterizer_impl.h).
Outputs go to oracle/_ref/<variant>/ (git-ignored, but shipped to the GPU box by gpurun).

The resulting modules are the reference's pybind ``_C`` modules:
  ref_fov_C : rasterize_gaussians(24 args), mark_visible      (FOV/rasterize_points.h:17-44, FOV/ext.cpp:15-18)
  ref_obb_C : rasterize_gaussians(19 args), rasterize_gaussians_backward, mark_visible
  ref_sum_C : rasterize_gaussians(19 args) -> 8-tuple, rasterize_gaussians_backward(21 args), mark_visible
  ref_max_C / ref_lwmc_C : as ref_sum_C (lwmc: 20 args, loss_map before debug)
  ref_naive_C / ref_mmfr_C : the SMFR / MMFR foveation baselines (23 args; mmfr takes cur_level instead of highest_levels)
  ref_vanilla_C : the stock diff-gaussian-rasterization (19 args -> 6-tuple, rasterize_gaussians_backward(19 args), mark_visible)

Nothing in the product path imports these.  Only tests/, bench.py --impl reference and tools/make_golden.py do.
"""
import os
import sys
import glob
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("FOVGS_REFERENCE_ROOT", "/root/reference")
SUB = os.path.join(REF_ROOT, "fov3dgs", "submodules")

VARIANTS = {
    # name -> (directory under submodules/, has backward.cu)
    "ref_fov_C": ("diff-gaussian-rasterization_fov_pcheck_obb", False),
    "ref_obb_C": ("diff-gaussian-rasterization_pcheck_obb", True),
    "ref_sum_C": ("diff-gaussian-rasterization_pcheck_obb_sum", True),
    # SURVEY.md §8f "next" rows: pruning-metric variants and the two foveation baselines
    "ref_max_C": ("diff-gaussian-rasterization_pcheck_obb_max", True),
    "ref_lwmc_C": ("diff-gaussian-rasterization_pcheck_obb_loss_weighted_max_count", True),
    "ref_naive_C": ("diff-gaussian-rasterization_naive_pcheck_obb", False),
    "ref_mmfr_C": ("diff-gaussian-rasterization_mmfr_pcheck_obb", False),
    # the stock Inria rasterizer the reference vendors (fov3dgs/gaussian_wrapper.py:2 cuda_type="original")
    "ref_vanilla_C": ("diff-gaussian-rasterization", True),
}


def so_path(name):
    hits = glob.glob(os.path.join(OUT, name, name + "*.so"))
    return hits[0] if hits else None


def build_one(name, verbose=False):
    """Compile one reference variant for sm_100 into oracle/_ref/<name>/. Returns the .so path."""
    sub, has_bwd = VARIANTS[name]
    root = os.path.join(SUB, sub)
    if not os.path.isdir(root):
        raise FileNotFoundError(f"reference sources not found at {root}")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    srcs = [
        os.path.join(root, "cuda_rasterizer", "rasterizer_impl.cu"),
        os.path.join(root, "cuda_rasterizer", "forward.cu"),
        os.path.join(root, "rasterize_points.cu"),
        os.path.join(root, "ext.cpp"),
    ]
    if has_bwd:
        srcs.insert(2, os.path.join(root, "cuda_rasterizer", "backward.cu"))
    bdir = os.path.join(OUT, name)
    os.makedirs(bdir, exist_ok=True)
    load(
        name=name,
        sources=srcs,
        extra_include_paths=[os.path.join(root, "third_party", "glm")],
        extra_cuda_cflags=["-include", "cstdint"],
        extra_cflags=["-include", "cstdint"],
        build_directory=bdir,
        verbose=verbose,
        is_python_module=False,
    )
    # keep only the .so (objects are large and not needed on the GPU box)
    for f in glob.glob(os.path.join(bdir, "*.o")):
        os.remove(f)
    return so_path(name)


# the variants bench.py --impl reference and the full-size parity tests (tests/test_gpu_fullsize_reference.py) need:
# FOV (headline, config 3), OBB (config 2) and SUM (config 5).  __graft_entry__.build() builds only these (about 3 minutes
# each when missing) unless FOVGS_BUILD_ALL_REFS=1 — the other four (parity tools, goldens) take ~15 more minutes
BENCH_VARIANTS = ("ref_fov_C", "ref_obb_C", "ref_sum_C")


def build_all(force=False, verbose=False, names=None):
    built = {}
    for name in (names or VARIANTS):
        p = so_path(name)
        if p and not force:
            built[name] = p
            continue
        if not os.path.isdir(REF_ROOT):
            # GPU box: reference sources are absent, only prebuilt .so files can be used.
            built[name] = None
            continue
        if force and os.path.isdir(os.path.join(OUT, name)):
            shutil.rmtree(os.path.join(OUT, name))
        built[name] = build_one(name, verbose=verbose)
    if built.get("ref_fov_C"):
        # machine code of the reference's two tile-level kernels, for oracle/ref_tile_tables.py (cuobjdump -xelf of the .so above)
        import ref_tile_tables
        ref_tile_tables.extract_cubin("ref_fov_C", force=force)
    return built


def load_ref(name):
    """Import a prebuilt reference module (returns the pybind module) or None if unavailable."""
    p = so_path(name)
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (the .so links libtorch)

    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    force = "--force" in sys.argv
    res = build_all(force=force, verbose="-v" in sys.argv)
    for k, v in res.items():
        print(k, "->", v)
