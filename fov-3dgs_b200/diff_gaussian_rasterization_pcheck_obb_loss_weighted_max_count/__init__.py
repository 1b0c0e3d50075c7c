"""Import-only stand-in: fov3dgs/gaussian_wrapper.py:2-7 imports this package at module import time, but it is a
pruning-metric / vanilla variant outside the hot path of this round (SURVEY.md §8f "next")."""
from fovgs.surface import make_unavailable_api as _make

globals().update(_make("diff_gaussian_rasterization_pcheck_obb_loss_weighted_max_count", "pruning-metric / vanilla variants are scheduled after the hot path (SURVEY.md section 8f)"))
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
