"""Drop-in for the reference package of the same name — the SMFR ("naive FR") foveation baseline: the foveated pipeline
with one shared model, levels only subset the Gaussians (reference: fov3dgs/submodules/
diff-gaussian-rasterization_naive_pcheck_obb/diff_gaussian_rasterization_naive_pcheck_obb/__init__.py; used by
fov3dgs/gaussian_renderer_fov_naive/__init__.py and render_compose_gazes_fps_naive.py)."""
from fovgs.surface import make_smfr_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
