"""Drop-in for the reference package of the same name (PS=1 pruning metric "max": forward with per-(pixel, Gaussian) hit
counts and the maximum alpha*T per Gaussian, SUM's backward; reference:
fov3dgs/submodules/diff-gaussian-rasterization_pcheck_obb_max/diff_gaussian_rasterization_pcheck_obb_max/__init__.py,
cuda_rasterizer/forward.cu:381,400)."""
from fovgs.surface import make_max_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
