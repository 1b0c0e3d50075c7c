"""Drop-in for the reference package of the same name — the MMFR foveation baseline: one rasterizer call per level model
(`cur_level`), each rendering only its level's tiles; the caller adds the four images (reference: fov3dgs/submodules/
diff-gaussian-rasterization_mmfr_pcheck_obb/diff_gaussian_rasterization_mmfr_pcheck_obb/__init__.py; used by
fov3dgs/gaussian_renderer_fov_mmfr/__init__.py and render_compose_gazes_fps_mmfr.py)."""
from fovgs.surface import make_mmfr_api as _make

globals().update(_make())
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
