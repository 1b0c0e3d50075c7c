"""Drop-in for the reference package of the same name (PS=1 training: forward + per-Gaussian counters + backward;
reference: fov3dgs/submodules/diff-gaussian-rasterization_pcheck_obb_sum/diff_gaussian_rasterization_pcheck_obb_sum/__init__.py)."""
from fovgs.surface import make_sum_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
