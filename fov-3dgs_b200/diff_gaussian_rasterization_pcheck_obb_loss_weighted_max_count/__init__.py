"""Drop-in for the reference package of the same name (PS=1 pruning metric "CE": every pixel adds loss_map[pixel] to the
Gaussian with its largest alpha*T; extra `loss_map` argument; SUM's backward; reference:
fov3dgs/submodules/diff-gaussian-rasterization_pcheck_obb_loss_weighted_max_count/
diff_gaussian_rasterization_pcheck_obb_loss_weighted_max_count/__init__.py, cuda_rasterizer/forward.cu:403-410,435)."""
from fovgs.surface import make_lwmc_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
