"""Drop-in for the reference package of the same name (foveated forward; reference:
fov3dgs/submodules/diff-gaussian-rasterization_fov_pcheck_obb/diff_gaussian_rasterization_fov_pcheck_obb/__init__.py)."""
from fovgs.surface import make_fov_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
