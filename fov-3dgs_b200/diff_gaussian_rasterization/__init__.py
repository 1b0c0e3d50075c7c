"""Drop-in for the stock Inria rasterizer the reference vendors (fov3dgs/gaussian_wrapper.py:2,11 cuda_type="original";
reference: fov3dgs/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py): forward + backward,
no OBB tile test, no -4.5 falloff cut, returns (color, radii)."""
from fovgs.surface import make_vanilla_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
